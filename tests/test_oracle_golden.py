"""CPU tests: the oracle (oracle/*.c, oracle/port.py) against the golden vectors that were produced
by running the unmodified reference (tests/golden/make_golden.py), plus - where /root/reference is
present - live pinning against the reference itself."""
import os

import numpy as np
import pytest
import torch

from oracle import port, refine_oracle as ro, ref_harness as rh
from boxfusion_b200.driver import FusionSession
from boxfusion_b200.synthetic import SyntheticScene, make_cfg
from tests.golden.make_golden import SEQUENCES


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8) if a.dtype.kind == "f" else a


def test_sampled_iou_c_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "iou_pairs.npz"))
    gate, cnt = port.obb_counts_pairs_c(g["corners"], g["ia"], g["ib"])
    iou = np.where(gate > 0, port.iou_from_counts(cnt), 0.0)
    assert np.array_equal(iou, g["iou"])                 # float64, exact
    n = len(g["ov_iou"])
    cc = np.concatenate([g["ov_a"], g["ov_b"]])
    gate, cnt = port.obb_counts_pairs_c(cc, np.arange(n), np.arange(n) + n)
    assert np.array_equal(np.where(gate > 0, port.iou_from_counts(cnt), 0.0), g["ov_iou"])


def test_sampled_iou_scipy_port_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "iou_pairs.npz"))
    for i in range(0, 60):
        gate, cnt = port.obb_counts_scipy(g["ov_a"][i], g["ov_b"][i])
        assert (float(port.iou_from_counts(cnt)) if gate else 0.0) == g["ov_iou"][i]


def test_hull_planes_match_qhull():
    from scipy.spatial import ConvexHull
    from boxfusion_b200.synthetic import random_boxes
    t, R = random_boxes(200, 7, tilt_noise=0.02)
    c = port.GeneralInstance3DBoxes(torch.from_numpy(t), torch.from_numpy(R)).corners.numpy()
    for n in range(200):
        eq = ConvexHull(c[n]).equations
        mine = port.hull_planes_c(c[n])
        d = np.abs(eq[:, None, :] - mine[None, :, :]).max(-1)
        assert d.min(1).max() < 1e-13 and d.min(0).max() < 1e-13


def test_corners_match_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "iou_pairs.npz"))
    c = port.GeneralInstance3DBoxes(torch.from_numpy(g["tensor"]), torch.from_numpy(g["R"])).corners.numpy()
    assert np.array_equal(_bits(c), _bits(g["corners"]))


@pytest.mark.parametrize("V", [3, 5, 8])
def test_refine_oracle_matches_reference_golden(golden_dir, V):
    g = np.load(os.path.join(golden_dir, "refine_cases.npz"))
    pst = np.load(os.path.join(golden_dir, "pst_1024_0.npy"))
    cfg = make_cfg("ca1m", pst_size=1024)
    W, H = g[f"v{V}_size"]
    K16 = ro.K16_from_K3(g[f"v{V}_K"])
    T, R, S, P, proj = (g[f"v{V}_{k}"] for k in ("tensor", "R", "scores", "poses", "projected"))
    search = np.array([0.1, 0.1, 0.1, 0.5, 0.5, 0.5], np.float32)
    cs = ro.make_cfg_struct(cfg, H, W)
    ro.lib().bfo_reset_stats()
    for b in range(T.shape[0]):
        fit = ro.evaluate(T[b, 0], proj[b], pst, R[b, 0], P[b], K16, search, H, W)
        assert np.array_equal(_bits(fit), _bits(g[f"v{V}_fitness0"][b]))
        upd, out6, n_it, _ = ro.refine_box(T[b], R[b], S[b], proj[b], P[b], pst, K16, cs)
        assert upd == bool(g[f"v{V}_flag"][b])
        if upd:
            assert np.array_equal(_bits(out6), _bits(g[f"v{V}_fused"][b]))


def _replay(impl, name, golden_dir, frames=None):
    spec = dict(SEQUENCES[name])
    n_frames = spec.pop("frames")
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    scene = SyntheticScene(**spec)
    cfg = make_cfg(spec["shape"], pst_path=os.path.join(golden_dir, "pst_1024_0.npy"), pst_size=1024)
    sess = FusionSession(impl, cfg)
    for k in range(frames or n_frames):
        kf = scene.keyframe(k)
        ins, pose_np = sess.pred_instances_from_world(kf, g[f"k{k}_tensor_w"], g[f"k{k}_R_w"], g[f"k{k}_projected"])
        sess.step(kf, ins, pose_np)
        snap = sess.snapshot()
        for key, val in snap.items():
            ref = g[f"k{k}_snap_{key}"]
            assert val.shape == ref.shape and np.array_equal(_bits(val), _bits(ref)), (name, k, key)


@pytest.mark.parametrize("name", list(SEQUENCES))
def test_port_sequence_matches_reference_golden(golden_dir, name, monkeypatch):
    monkeypatch.setattr(port, "IOU_BACKEND", "c")
    _replay(port, name, golden_dir)


def test_port_sequence_scipy_backend(golden_dir, monkeypatch):
    monkeypatch.setattr(port, "IOU_BACKEND", "scipy")
    _replay(port, "seq_ca1m", golden_dir, frames=6)


def test_lift_and_project_match_golden(golden_dir):
    """demo.py:220-221 through the port's containers reproduces the reference's tensors."""
    name = "seq_scannet_tilt"
    spec = dict(SEQUENCES[name]); spec.pop("frames")
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    scene = SyntheticScene(**spec)
    cfg = make_cfg(spec["shape"], pst_path=os.path.join(golden_dir, "pst_1024_0.npy"))
    sess = FusionSession(port, cfg)
    for k in range(3):
        ins, _ = sess.make_pred_instances(scene.keyframe(k))
        assert np.array_equal(_bits(ins.pred_boxes_3d.tensor.numpy()), _bits(g[f"k{k}_tensor_w"]))
        assert np.array_equal(_bits(ins.projected_boxes.numpy()), _bits(g[f"k{k}_projected"]))


# ---- round 2 pins: detection pre-filters, pose disparity, result formats, frame-unit gap -----------------------------

def test_port_prefilters_match_reference_golden(golden_dir):
    """oracle/port.py's check_uv_bounds / check_floor_mask / check_large_mask == the reference's (box_manager.py:217-245)."""
    g = np.load(os.path.join(golden_dir, "prefilters.npz"))
    tt, uu = torch.from_numpy(g["tensor"]), torch.from_numpy(g["uv"])
    for shape in ("ca1m", "scannet"):
        cfg = make_cfg(shape)
        bm = port.BoxManager(cfg)
        W, H = cfg["cam"]["W"], cfg["cam"]["H"]
        for ratio in (cfg["detection"]["uv_bound_value"], 1.0, 0.75):
            assert np.array_equal(bm.check_uv_bounds(uu, W, H, ratio=ratio).numpy(), g[f"{shape}_uv_{ratio}"])
        for ratio in (cfg["detection"]["floor_ratio"], 20):
            assert np.array_equal(bm.check_floor_mask(tt, ratio=ratio).numpy(), g[f"{shape}_floor_{ratio}"])
        for thres in (0.5, 2.5):
            assert np.array_equal(bm.check_large_mask(tt, thres=thres).numpy(), g[f"{shape}_large_{thres}"])


def test_port_pose_disparity_matches_reference_golden(golden_dir):
    """compute_pose_disparity / compute_pose_center_disparity (box_manager.py:168-215): bit-identical (same torch ops)."""
    g = np.load(os.path.join(golden_dir, "prefilters.npz"))
    bm = port.BoxManager(make_cfg("ca1m"))
    P = torch.from_numpy(g["poses"])
    for k in range(0, len(g["ia"]), 7):
        a, b = int(g["ia"][k]), int(g["ib"][k])
        base, ang, score, cd = bm.compute_pose_center_disparity(P[a], P[b], g["centers"][a], g["centers"][b])
        got = np.array([float(base), float(ang), float(score), float(cd)])
        assert np.array_equal(got, g["disparity"][k], equal_nan=True), k


def test_port_sequence_with_frame_gap_matches_reference_golden(golden_dir):
    """A keyframe every 3rd frame, check_valid_num on: frame ids and the `count - gap` threshold are in FRAMES."""
    from tests.golden.make_golden import GAP_SEQUENCE
    name, spec, gap = GAP_SEQUENCE
    spec = dict(spec)
    n_frames = spec.pop("frames")
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    port.IOU_BACKEND = "c"
    cfg = make_cfg(spec["shape"], pst_path=os.path.join(golden_dir, "pst_1024_0.npy"), pst_size=1024)
    cfg["data"]["gap"] = gap
    cfg["box_fusion"]["check_valid"] = True
    scene = SyntheticScene(**spec)
    sess = FusionSession(port, cfg, frame_stride=gap)
    dropped = 0
    for k in range(n_frames):
        sess.step(scene.keyframe(k))
        snap = sess.snapshot()
        for key, val in snap.items():
            ref = g[f"k{k}_snap_{key}"]
            assert val.shape == ref.shape and np.array_equal(_bits(val), _bits(ref)), (k, key)
        if sess.last_keep_idx is not None:
            dropped += max(0, len(sess.last_keep_idx) - len(sess.all_pred_box))
    assert dropped > 0, "the golden sequence must exercise check_valid_num"


def test_result_formats_host_part_matches_reference_golden(golden_dir):
    """results.post_process == the reference's tools/utils.py post_process (:302-317) on the golden corner arrays."""
    from boxfusion_b200 import results
    g = np.load(os.path.join(golden_dir, "results_formats.npz"))
    assert np.array_equal(results.post_process(g["corners"]), g["kept"])
    assert np.array_equal(results.post_process(g["corners"], threshold=0.5), g["kept05"])
    assert np.array_equal(results.post_process(torch.from_numpy(g["corners"])).numpy(), g["kept"])


# ---- live pinning (build container only) ------------------------------------------------------

needs_ref = pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")


@needs_ref
def test_live_kernel_string_vs_c_oracle():
    """oracle/refine_oracle.c == the reference's kernel string compiled for the host, bit for bit."""
    from boxfusion_b200.synthetic import refine_problem, make_pst
    ref = rh.load_reference()
    pst = make_pst(512, seed=3)
    import cv2, tempfile
    path = os.path.join(tempfile.mkdtemp(), "pst512.tiff")
    cv2.imwrite(path, pst)
    cfg = make_cfg("scannet", pst_path=path, pst_size=512)
    bf = ref.BoxFusion(cfg)
    prob = refine_problem(4, 6, seed=9, shape="scannet")
    W, H = prob["size"]
    bf.update_intrinsics((W, H), prob["K"])
    K16 = ro.K16_from_K3(prob["K"])
    rs = np.random.RandomState(0)
    for b in range(4):
        # observation corners: reuse the other views' true projections via the port's projector
        ins = port.Instances3D((H, W))
        ins.pred_boxes_3d = port.GeneralInstance3DBoxes(torch.from_numpy(prob["tensor"][b]), torch.from_numpy(prob["R"][b]))
        ins.cam_pose = torch.from_numpy(prob["poses"][b])
        ins.project_3d_boxes(prob["K"], H=H, W=W)
        proj = ins.projected_boxes.numpy()
        search = rs.uniform(0.01, 0.5, 6).astype(np.float32)
        f_ref = bf.evaluate_iou(prob["tensor"][b, 0].astype(np.float64), proj, prob["R"][b, 0], prob["scores"][b],
                                prob["poses"][b], search, 6)
        f_me = ro.evaluate(prob["tensor"][b, 0], proj, pst, prob["R"][b, 0], prob["poses"][b], K16, search, H, W, 512)
        assert np.array_equal(_bits(f_ref), _bits(f_me))


@needs_ref
def test_live_reference_vs_port_sequence(monkeypatch):
    monkeypatch.setattr(port, "IOU_BACKEND", "c")
    ref = rh.load_reference()
    scene = SyntheticScene(n_objects=30, seed=5, max_det=15, shape="ca1m", tilt_noise=0.01)
    cfg = make_cfg("ca1m", pst_path=os.path.join(rh.REFERENCE_ROOT, "data", "pst_1024_0.tiff"), pst_size=1024)
    a, b = FusionSession(ref, cfg), FusionSession(port, cfg)
    for k in range(6):
        kf = scene.keyframe(k)
        a.step(kf); b.step(kf)
        sa, sb = a.snapshot(), b.snapshot()
        for key in sa:
            assert sa[key].shape == sb[key].shape and np.array_equal(_bits(sa[key]), _bits(sb[key])), (k, key)


@needs_ref
def test_live_reference_vs_port_check_valid(monkeypatch):
    """BoxManager.check_valid_num (box_manager.py:151-166, cfg box_fusion.check_valid; off in the shipped configs): the
    unmodified reference against the port on a scene where boxes are never re-observed and get dropped."""
    monkeypatch.setattr(port, "IOU_BACKEND", "c")
    ref = rh.load_reference()
    scene = SyntheticScene(n_objects=40, seed=21, max_det=12, shape="scannet", new_frac=0.35)
    cfg = make_cfg("scannet", pst_path=os.path.join(rh.REFERENCE_ROOT, "data", "pst_1024_0.tiff"), pst_size=1024)
    cfg["box_fusion"]["check_valid"] = True
    cfg["data"]["gap"] = 2
    a, b = FusionSession(ref, cfg), FusionSession(port, cfg)
    dropped = 0
    for k in range(8):
        kf = scene.keyframe(k)
        a.step(kf); b.step(kf)
        sa, sb = a.snapshot(), b.snapshot()
        for key in sa:
            assert sa[key].shape == sb[key].shape and np.array_equal(_bits(sa[key]), _bits(sb[key])), (k, key)
        if a.last_keep_idx is not None:
            dropped += max(0, len(a.last_keep_idx) - len(a.all_pred_box))
    assert dropped > 0
