"""CPU parity of the kernel's evaluation core (boxfusion_b200/csrc/bf_refine_eval.cuh) against the oracle.

The header is `__host__`-compilable on purpose: tests/host_eval/bf_eval_host.cpp wraps the very source nvcc compiles
into bf_refine_kernel, built here with g++ -ffp-contract=off (the counterpart of nvcc -fmad=false).  The certified
side classification (which of the reference's ray casts / float64 segment tests can be skipped) must never change a
result: every float32 fitness / IoU is compared BIT FOR BIT with the oracle (pinned to the reference,
tests/test_oracle_golden.py), on generic inputs and on inputs built to sit on the decision boundaries (integer grids,
clamped onto the image border, identical polygons, shared vertices, sub-pixel polygons).
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from boxfusion_b200.synthetic import make_pst, refine_problem
from oracle import port, refine_oracle as ro

HERE = os.path.dirname(os.path.abspath(__file__))
FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def libs():
    src = os.path.join(HERE, "host_eval", "bf_eval_host.cpp")
    hdr = os.path.join(os.path.dirname(HERE), "boxfusion_b200", "csrc", "bf_refine_eval.cuh")
    out_dir = os.path.join(os.path.dirname(HERE), "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libbf_eval_host.so")
    if not os.path.isfile(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-fopenmp",
                        "-Wno-unknown-pragmas", "-o", out, src, "-lm"], check=True)
    lh = ctypes.CDLL(out)
    lh.bfh_iou_points.restype = ctypes.c_float
    lh.bfh_iou_points.argtypes = [FP, FP, ctypes.c_float, ctypes.c_float, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lh.bfh_evaluate.restype = None
    lh.bfh_evaluate.argtypes = ([FP, FP, FP, ctypes.c_int, ctypes.c_int, FP, FP, ctypes.c_int] + [ctypes.c_float] * 4 +
                                [FP, ctypes.c_float, ctypes.c_float, FP, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int])
    lo = ro.lib()
    lo.bfo_iou_points.restype = ctypes.c_float
    lo.bfo_iou_points.argtypes = [FP, FP]
    return lh, lo


W, H = 384.0, 512.0


def _gen(kind, rs):
    if kind == "generic":
        c = rs.uniform([50, 50], [W - 50, H - 50]); a = c + rs.normal(0, 60, (8, 2)); b = c + rs.normal(0, 60, (8, 2)) + rs.normal(0, 20, 2)
    elif kind == "integer_grid":
        c = rs.randint(5, 30, 2); a = c + rs.randint(-6, 7, (8, 2)); b = c + rs.randint(-6, 7, (8, 2))
    elif kind == "clamped_to_border":
        c = rs.uniform([-50, -50], [W + 50, H + 50]); a = c + rs.normal(0, 80, (8, 2)); b = a + rs.normal(0, 10, (8, 2))
        a = np.clip(a, [0, 0], [W, H]); b = np.clip(b, [0, 0], [W, H])
    elif kind == "identical":
        c = rs.uniform([50, 50], [W - 50, H - 50]); a = c + rs.normal(0, 60, (8, 2)); b = a.copy()
        if rs.rand() < 0.5:
            b = b[rs.permutation(8)]
        if rs.rand() < 0.5:
            b[rs.randint(8)] += rs.normal(0, 1e-3, 2)
    elif kind == "shared_vertices":
        c = rs.randint(5, 30, 2) * 8.0; a = c + rs.randint(-6, 7, (8, 2)) * 4.0; b = a.copy(); k = rs.randint(1, 8)
        b[:k] = c + rs.randint(-6, 7, (k, 2)) * 4.0
    elif kind == "sub_pixel":
        c = rs.uniform([50, 50], [W - 50, H - 50]); a = c + rs.normal(0, 0.01, (8, 2)); b = c + rs.normal(0, 0.01, (8, 2))
    elif kind == "nearly_identical":   # shallow crossings, vertices within the classification margin of the other boundary
        c = rs.uniform([50, 50], [W - 50, H - 50]); base = rs.normal(0, 40, (3, 2))
        sg = np.array([[i & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)]) - 0.5
        eps = 10.0 ** rs.uniform(-4, 0)
        a = c + sg @ base; b = c + sg @ (base + rs.normal(0, eps, (3, 2))) + rs.normal(0, eps, 2)
    elif kind == "tiny_rotation":
        c = rs.uniform([80, 80], [W - 80, H - 80]); base = rs.normal(0, 30, (3, 2))
        sg = np.array([[i & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)]) - 0.5
        th = 10.0 ** rs.uniform(-6, -1); R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        a = c + sg @ base; b = c + (sg @ base) @ R.T * (1 + rs.normal(0, 1e-4))
    else:   # "box_like": two projected parallelepipeds, slightly different
        c = rs.uniform([50, 50], [W - 50, H - 50]); base = rs.normal(0, 40, (3, 2))
        sg = np.array([[i & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)]) - 0.5
        a = c + sg @ base; b = c + sg @ (base + rs.normal(0, 3, (3, 2))) + rs.normal(0, 3, 2)
    return np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)


@pytest.mark.parametrize("rolled", [0, 1])          # the kernel's two code-size variants of the same evaluation
@pytest.mark.parametrize("kind", ["generic", "integer_grid", "clamped_to_border", "identical", "shared_vertices",
                                  "sub_pixel", "box_like", "nearly_identical", "tiny_rotation"])
def test_polygon_iou_bit_exact(libs, kind, rolled):
    lh, lo = libs
    rs = np.random.RandomState(abs(hash(kind)) % (2 ** 31))
    fb, ov = ctypes.c_int(0), ctypes.c_int(0)
    n_fb = 0
    for _ in range(12000):
        a, b = _gen(kind, rs)
        r = lo.bfo_iou_points(a.ctypes.data_as(FP), b.ctypes.data_as(FP))
        g = lh.bfh_iou_points(a.ctypes.data_as(FP), b.ctypes.data_as(FP), W, H, ctypes.byref(fb), ctypes.byref(ov), rolled)
        assert ov.value == 0
        assert np.float32(r).view(np.uint32) == np.float32(g).view(np.uint32), (kind, r, g, a.tolist(), b.tolist())
        n_fb += fb.value > 0
    if kind == "generic":
        assert n_fb < 240          # the certified classification decides (almost) everything away from degeneracy


@pytest.mark.parametrize("rolled", [0, 1])
@pytest.mark.parametrize("B,V,P,shape,scale,off", [(6, 6, 1024, "ca1m", 1.0, 0.0), (6, 8, 512, "scannet", 1.0, 0.0),
                                                   (3, 32, 1024, "ca1m", 1.0, 0.0), (6, 6, 1024, "ca1m", 4.0, 0.0),
                                                   # cameras looking past the object: projections clamped onto image borders
                                                   # and corners (the certified shared-border rule, bf_border_of)
                                                   (6, 8, 512, "scannet", 1.0, 0.9), (6, 6, 1024, "ca1m", 4.0, 1.2)])
def test_fitness_bit_exact(libs, B, V, P, shape, scale, off, rolled):
    """bf_evaluate_kernel's loops on the host == oracle evaluate (box_fusion.py:413-461), every particle, every bit."""
    lh, _ = libs
    prob = refine_problem(B, V, seed=B * 7 + V, shape=shape)
    if off > 0:
        from boxfusion_b200.synthetic import look_at_pose
        rs0 = np.random.RandomState(77 + V)
        for b in range(B):
            c = prob["tensor"][b, :, :3].mean(0)
            for v in range(V):
                d = rs0.normal(0, 1, 3)
                prob["poses"][b, v] = look_at_pose(prob["poses"][b, v][:3, 3].astype(np.float64),
                                                   c + off * d / np.linalg.norm(d)).astype(np.float32)
    Wi, Hi = prob["size"]
    pst = make_pst(P, seed=1)
    K16 = ro.K16_from_K3(prob["K"])
    rs = np.random.RandomState(V)
    stats = (ctypes.c_longlong * 4)()

    def f(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        return a, a.ctypes.data_as(FP)

    for b in range(B):
        t, R, po = prob["tensor"][b], prob["R"][b], prob["poses"][b]
        ins = port.Instances3D((Hi, Wi))
        ins.pred_boxes_3d = port.GeneralInstance3DBoxes(torch.from_numpy(t), torch.from_numpy(R))
        ins.cam_pose = torch.from_numpy(po)
        ins.project_3d_boxes(prob["K"], H=Hi, W=Wi)
        uv = ins.projected_boxes.numpy().reshape(V, 16)
        search = (rs.uniform(0.005, 0.5, 6) * scale).astype(np.float32)
        box = (t[rs.randint(V)] + rs.normal(0, 0.03, 6)).astype(np.float32)
        ref = ro.evaluate(box, uv, pst, R[0], po, K16, search, Hi, Wi, P)
        out = np.zeros(P, np.float32)
        keep = [f(box), f(uv), f(pst), f(R[0].reshape(9)), f(po.reshape(V, 16)), f(search)]
        lh.bfh_evaluate(keep[0][1], keep[1][1], keep[2][1], P, P, keep[3][1], keep[4][1], V,
                        float(K16[0]), float(K16[2]), float(K16[5]), float(K16[6]), keep[5][1], float(Hi), float(Wi),
                        out.ctypes.data_as(FP), stats, rolled)
        assert np.array_equal(ref.view(np.uint32), out.view(np.uint32))
    assert stats[3] == 0
    if off == 0:
        assert stats[1] < 0.05 * stats[0]      # exact fallback tests are the exception (measured: ~1 % of evaluations)
