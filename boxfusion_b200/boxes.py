"""`boxfusion.boxes.GeneralInstance3DBoxes` for the B200 path (reference: boxfusion/boxes.py:656-943).

Only the box container the fusion hot path touches is provided (the MMDet3D-derived
classes in the first 650 lines of the reference file are detector-side and out of scope).  Geometry is
computed by the CUDA library: `corners` (bf_box_corners) and `transform2world` (bf_transform2world).
"""
from __future__ import annotations

import torch

from . import ops


class GeneralInstance3DBoxes(object):
    """tensor[N,6] = (x,y,z,l,h,w) and R[N,3,3]; same constructor/fields as the reference (boxes.py:657-669)."""

    def __init__(self, xyzlhw, R, box_dim=6 + 3 * 3, origin=(0.5, 0.5, 0), dof=None):
        device = xyzlhw.device if isinstance(xyzlhw, torch.Tensor) else torch.device("cpu")
        self.dof = dof
        self.box_dim = box_dim
        self.tensor = torch.as_tensor(xyzlhw, dtype=torch.float32, device=device).clone()
        self.R = torch.as_tensor(R, dtype=torch.float32, device=device).clone()

    @classmethod
    def _wrap(cls, tensor, R, dof=None):
        """Adopt freshly produced tensors (results of cat / indexing) without the defensive clone of __init__."""
        obj = cls.__new__(cls)
        obj.dof, obj.box_dim, obj.tensor, obj.R = dof, 6 + 3 * 3, tensor, R
        return obj

    @classmethod
    def empty(cls, dof=None):
        return cls(torch.zeros((0, 6)), torch.zeros((0, 3, 3)), dof=dof)

    # ---- cheap views (boxes.py:678-723) ----------------------------------------------------------
    @property
    def volume(self):
        return self.tensor[:, 3] * self.tensor[:, 4] * self.tensor[:, 5]

    @property
    def dims(self):
        return self.tensor[:, 3:6]

    @property
    def whl(self):
        return self.tensor[:, [5, 4, 3]]

    @property
    def xyzwhl(self):
        return self.tensor[:, [0, 1, 2, 5, 4, 3]]

    @property
    def gravity_center(self):
        return self.tensor[:, :3]

    center = gravity_center

    @property
    def device(self):
        return self.tensor.device

    # ---- geometry on the GPU ----------------------------------------------------------------------
    @property
    def corners(self):
        """[N,8,3] corners, vertex order of boxes.py:737-766; returned on the tensor's own device."""
        if len(self) == 0:
            return torch.zeros((0, 8, 3), dtype=torch.float32, device=self.tensor.device)
        c = ops.box_corners(self.tensor, self.R)
        return c if self.tensor.is_cuda else c.to(self.tensor.device)

    def transform2world(self, cam_pose):
        """boxes.py:825-833: centre <- R_c c + t_c, R <- R_c R, in place."""
        if not isinstance(cam_pose, torch.Tensor):
            cam_pose = torch.from_numpy(cam_pose)
        if len(self) == 0:
            return
        if self.tensor.is_cuda:
            t, r = self.tensor.contiguous(), self.R.contiguous()
            one = ops.shared_pose(cam_pose) if cam_pose.shape[0] == t.shape[0] else None
            if one is not None:                                  # demo.py:216: one pose for the keyframe's detections
                ops.transform2world_pose_(t, r.view(-1, 9) if r.dim() == 3 else r, one)
            else:
                ops.transform2world_(t, r, cam_pose)
            self.tensor, self.R = t, r
        else:
            dev = ops._dev()
            t, r = self.tensor.to(dev).contiguous(), self.R.to(dev).contiguous()
            ops.transform2world_(t, r, cam_pose)
            self.tensor.copy_(t)
            self.R = r.to(self.R.device)

    def translate(self, trans_vector):
        if not isinstance(trans_vector, torch.Tensor):
            trans_vector = self.tensor.new_tensor(trans_vector)
        self.tensor[:, :3] += trans_vector

    # ---- container protocol (boxes.py:845-943) ----------------------------------------------------
    def __getitem__(self, item):
        if isinstance(item, int):
            return type(self)(self.tensor[item].view(1, -1), self.R[item].view(1, 3, 3), dof=self.dof)
        b, r = self.tensor[item], self.R[item]
        assert b.dim() == 2, f"Indexing on Boxes with {item} failed to return a matrix!"
        if isinstance(item, slice):
            return type(self)(b, r, dof=self.dof)              # a slice is a view: clone like the reference
        return type(self)._wrap(b, r, dof=self.dof)            # advanced indexing already copied

    def __len__(self):
        return self.tensor.shape[0]

    def __repr__(self):
        return self.__class__.__name__ + "(\n    " + str(self.tensor) + ")"

    @classmethod
    def cat(cls, boxes_list):
        assert isinstance(boxes_list, (list, tuple))
        if len(boxes_list) == 0:
            return cls.empty()
        assert all(isinstance(b, cls) for b in boxes_list)
        return cls._wrap(torch.cat([b.tensor for b in boxes_list], dim=0), torch.cat([b.R for b in boxes_list], dim=0),
                         dof=boxes_list[0].dof)

    def split(self, split_size_or_sections):
        return [type(self)(t, r, dof=self.dof) for t, r in
                zip(torch.split(self.tensor, split_size_or_sections), torch.split(self.R, split_size_or_sections))]

    def to(self, device):
        return type(self)(self.tensor.to(device), self.R.to(device), dof=self.dof)

    def clone(self):
        return type(self)(self.tensor.clone(), self.R.clone(), dof=self.dof)

    def __iter__(self):
        yield from self.tensor
