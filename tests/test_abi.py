"""CPU test: the C-ABI library builds, loads, and exports every symbol include/boxfusion_b200.h declares;
the product refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import pytest
import torch

from boxfusion_b200 import _lib, build as bf_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    bf_build.build()
    lib = _lib.load_library()
    header = open(os.path.join(ROOT, "include", "boxfusion_b200.h")).read()
    declared = set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", header)) - {"bf_handle"}
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.bf_version() == 100 and lib.bf_fusion_cap() == 32


def test_refine_cfg_layout_matches_header():
    assert ctypes_sizeof() == 72          # 17 x 4-byte fields + one double, padded to 8


def ctypes_sizeof():
    import ctypes
    return ctypes.sizeof(_lib.RefineCfg)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    import numpy as np
    from boxfusion_b200 import api, ops
    with pytest.raises(RuntimeError):
        ops.box_corners(np.zeros((2, 6), np.float32), np.zeros((2, 3, 3), np.float32))
    with pytest.raises(RuntimeError):
        api.Instances3D.obb_iou(np.zeros((8, 3), np.float32), np.zeros((8, 3), np.float32))
