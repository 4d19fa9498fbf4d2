"""CPU tests of the product's host-side logic (no GPU, no compute calls): the reference-shaped containers, BoxManager's
list bookkeeping and device layout, keyframe packing, recorder / player, result filters."""
import numpy as np
import pytest
import torch

from boxfusion_b200 import replay, results
from boxfusion_b200.box_manager import BoxManager
from boxfusion_b200.boxes import GeneralInstance3DBoxes
from boxfusion_b200.engine import pack_keyframe
from boxfusion_b200.instances import Instances3D
from boxfusion_b200.synthetic import SyntheticScene, make_cfg, make_pst
from oracle import port


def _inst(n, off=0):
    ins = Instances3D((480, 640))
    ins.scores = torch.arange(n, dtype=torch.float32) + off
    ins.pred_boxes_3d = GeneralInstance3DBoxes(torch.rand(n, 6), torch.eye(3).repeat(n, 1, 1))
    ins.categories = np.array([f"c{i + off}" for i in range(n)])
    ins.names = [f"n{i + off}" for i in range(n)]
    return ins


def test_instances3d_container_semantics_match_the_port():
    """Same field/indexing/cat behaviour as the reference container (instances.py:128-331), checked against the port."""
    a, b = _inst(4), _inst(3, off=10)
    c = Instances3D.cat([a, b])
    assert len(c) == 7 and c.names[4] == "n10" and c.categories[5] == "c11"
    assert isinstance(c.pred_boxes_3d, GeneralInstance3DBoxes) and c.pred_boxes_3d.tensor.shape == (7, 6)
    assert Instances3D.cat([a]) is a                                           # instances.py:313-314
    with pytest.raises(ValueError):
        c[np.array([6, 0, 2])]                                                 # list fields need a Bool/Long tensor (:249-261)
    c.remove("names")
    sub = c[np.array([6, 0, 2])]                                               # demo.py:325 / instances.py:440 index like this
    assert sub.scores.tolist() == [12.0, 0.0, 2.0] and sub.categories.tolist() == ["c12", "c0", "c2"]
    c.names = [f"n{i}" for i in range(4)] + [f"n{i + 10}" for i in range(3)]
    m = torch.tensor([True, False, True, False, False, False, True])
    assert c[m].names == ["n0", "n2", "n12"] and len(c[m]) == 3
    assert c[torch.tensor([1, 5])].names == ["n1", "n11"]
    assert len(c[2]) == 1 and c[2].scores.item() == 2.0 and len(c[1:4]) == 3
    with pytest.raises(IndexError):
        c[7]
    with pytest.raises(AssertionError):
        c.set("bad", torch.zeros(3))
    with pytest.raises(AttributeError):
        c.missing
    assert c.has("scores") and not c.has("nope") and "scores" in str(c)
    # the port (pinned to the reference) behaves the same on the same operations
    pa = port.Instances3D((480, 640)); pa.scores = a.scores.clone(); pa.names = list(a.names)
    pb = port.Instances3D((480, 640)); pb.scores = b.scores.clone(); pb.names = list(b.names)
    pc = port.Instances3D.cat([pa, pb])
    assert pc[m].names == c[m].names
    with pytest.raises(ValueError):
        pc[np.array([6, 0, 2])]


def test_boxes_container():
    b = GeneralInstance3DBoxes(torch.rand(5, 6), torch.eye(3).repeat(5, 1, 1))
    assert len(b) == 5 and b.dims.shape == (5, 3) and b[2].tensor.shape == (1, 6) and b[torch.tensor([0, 4])].R.shape == (2, 3, 3)
    c = GeneralInstance3DBoxes.cat([b, b[1:3]])
    assert len(c) == 7 and torch.equal(c.tensor[5], b.tensor[1])
    t0 = b.tensor.clone()
    b[1:3].tensor[:] = 0                                                       # slices are cloned, like the reference
    assert torch.equal(b.tensor, t0)
    assert torch.equal(b.volume, b.tensor[:, 3] * b.tensor[:, 4] * b.tensor[:, 5])


def test_box_manager_bookkeeping_and_device_layout():
    cfg = make_cfg("ca1m")
    bm = BoxManager(cfg)
    bm.init_new_predictions(3, 0)
    bm.init_new_predictions(2, 3)
    assert bm.fusion_list == [[0], [1], [2], [3], [4]] and bm.fusion_flag == [0] * 5
    bm.fusion_list[1] = [1, 7, 9]
    fl, ln, flag = bm.pack_lists(5)
    assert fl.shape == (5, 32) and ln.tolist() == [1, 3, 1, 1, 1] and fl[1, :3].tolist() == [1, 7, 9] and fl[1, 3:].sum() == 0
    fl2, ln2, flag2 = fl.copy(), ln.copy(), flag.copy()
    fl2[0, :2] = [0, 4]; ln2[0] = 2; flag2[3] = 1
    keep_ref = bm.fusion_list[1]
    bm.apply_lists(fl2, ln2, flag2, ln)
    assert bm.fusion_list[0] == [0, 4] and bm.fusion_list[1] is keep_ref and bm.fusion_flag == [0, 0, 0, 1, 0]
    bm.update(np.array([0, 1, 4]))
    assert bm.fusion_list == [[0, 4], [1, 7, 9], [4]] and len(bm.fusion_flag) == 5       # flags are not re-indexed (reference quirk)
    assert not bm.check_if_fusion([1, 7, 9])
    bm.add_fusion_ind(bm.fusion_list[1])
    bm.fusion_list[1].append(11)                                                        # deep copy was stored
    assert bm.check_if_fusion([1, 7, 9]) and bm.check_if_fusion([np.int64(1), 7, 9]) and not bm.check_if_fusion([1, 7, 9, 11])
    bm.already_fusion.append([5, 6, 8])                                                 # direct edits are picked up
    assert bm.check_if_fusion([5, 6, 8])
    bm.update_fusion_flag(0)
    assert bm.get_fusion_idx() == [0, 3] and bm.get_nofusion_idx() == [1, 2, 4]
    bm.fusion_list[0] = list(range(40))
    with pytest.raises(RuntimeError):
        bm.pack_lists(1)


def test_detection_masks_match_the_port():
    cfg = make_cfg("scannet")
    a, p = BoxManager(cfg), port.BoxManager(cfg)
    rs = np.random.RandomState(1)
    t = torch.from_numpy(np.concatenate([rs.normal(0, 2, (500, 3)), np.exp(rs.normal(-1, 1, (500, 3)))], 1).astype(np.float32))
    uv = torch.from_numpy(np.stack([rs.uniform(-30, 670, 500), rs.uniform(-30, 510, 500)], 1).astype(np.float32))
    assert torch.equal(a.check_uv_bounds(uv, 640, 480, ratio=0.9), p.check_uv_bounds(uv, 640, 480, ratio=0.9))
    assert torch.equal(a.check_floor_mask(t, ratio=15), p.check_floor_mask(t, ratio=15))
    assert torch.equal(a.check_large_mask(t, thres=1.5), p.check_large_mask(t, thres=1.5))


def test_pack_keyframe_layout_and_replay_roundtrip(tmp_path):
    kf = SyntheticScene(n_objects=20, seed=1, max_det=7).keyframe(2)
    n = kf.tensor_cam.shape[0]
    buf = pack_keyframe(kf.tensor_cam, kf.R_cam, kf.scores, kf.pred_boxes, kf.pred_proj_xy, kf.pose, kf.K, kf.image_size, frame_id=7)
    H = 56                                                     # BF_KF_HEADER (include/boxfusion_b200.h)
    assert buf.dtype == np.float32 and buf.shape == (H + 22 * n,)
    assert buf.view(np.int32)[0] == n and buf.view(np.int32)[1] == 7
    assert np.array_equal(buf[2:8], np.array([kf.K[0, 0], kf.K[1, 1], kf.K[0, 2], kf.K[1, 2], kf.image_size[0], kf.image_size[1]], np.float32))
    assert np.array_equal(buf[8:24].reshape(4, 4), kf.pose)
    assert np.array_equal(buf[24:40].reshape(4, 4), torch.linalg.inv(torch.from_numpy(kf.pose)).numpy())
    assert np.array_equal(buf[40:56].reshape(4, 4), np.linalg.inv(kf.pose))
    assert np.array_equal(buf[H + 6 * n:H + 15 * n].reshape(n, 3, 3), kf.R_cam) and np.array_equal(buf[H + 15 * n:H + 16 * n], kf.scores)
    rec = replay.KeyframeRecorder()
    rec.add_keyframe(kf)
    rec.add(5, kf.pose, kf.K, kf.image_size, kf.tensor_cam, kf.R_cam, torch.from_numpy(kf.scores), kf.pred_boxes, kf.pred_proj_xy)
    rec.save(str(tmp_path / "s.npz"))
    back = replay.load_sequence(str(tmp_path / "s.npz"))
    assert len(back) == 2 and back[1].frame_id == 5 and np.array_equal(back[0].R_cam, kf.R_cam) and back[0].image_size == kf.image_size


def test_post_process_filter():
    rs = np.random.RandomState(0)
    c = rs.uniform(0, 1, (50, 8, 3)).astype(np.float32) * rs.uniform(0.1, 1.0, (50, 1, 3)).astype(np.float32)
    ext = c.max(1) - c.min(1)
    assert np.array_equal(results.post_process(c, 0.3), c[(ext >= 0.3).all(1)])
    assert torch.equal(results.post_process(torch.from_numpy(c), 0.3), torch.from_numpy(c[(ext >= 0.3).all(1)]))


def test_small_public_helpers_match_reference_golden(golden_dir):
    """augment_vertices / IoU_2D / init_opt_params_v2 (instances.py:493-512, 616-641; box_fusion.py:602-619): host
    arithmetic mirrors kept for API completeness, against values produced by the unmodified reference
    (tests/golden/make_golden.py helpers)."""
    import os
    import numpy as np
    from boxfusion_b200 import api
    from boxfusion_b200.synthetic import make_cfg, make_pst
    g = np.load(os.path.join(golden_dir, "hull_helpers.npz"))
    for c, want in zip(g["aug_in"], g["aug"]):
        got = api.Instances3D.augment_vertices(c)
        assert got.dtype == want.dtype and np.array_equal(got, want)
    for A, B, want in zip(g["iou2d_A"], g["iou2d_B"], g["iou2d"]):
        iou, ov = api.Instances3D.IoU_2D(A, B)
        assert np.array_equal(iou, want[0]) and np.array_equal(ov, want[1])
    bf = api.BoxFusion(make_cfg("ca1m", pst_path=make_pst(64, seed=0), pst_size=64))
    for b, r, s_, m, rot in zip(g["v2_boxes"], g["v2_R"], g["v2_scores"], g["v2_mean"], g["v2_rot"]):
        mean, R = bf.init_opt_params_v2(b, r, s_)
        assert np.array_equal(mean, m) and np.array_equal(R, rot)
