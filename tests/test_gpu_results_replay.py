"""GPU tests of the section-8(f) rows 3 and 4: result formats and the keyframe recorder / player."""
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from boxfusion_b200 import api, replay, results                        # noqa: E402
from boxfusion_b200.driver import FusionSession                         # noqa: E402
from boxfusion_b200.engine import FusionEngine, pack_keyframe           # noqa: E402
from boxfusion_b200.synthetic import SyntheticScene, make_cfg, make_pst  # noqa: E402
from oracle import port                                                 # noqa: E402


def _session(n=5):
    scene = SyntheticScene(n_objects=25, seed=6, max_det=10, shape="scannet")
    cfg = make_cfg("scannet", pst_path=make_pst(128, seed=1), pst_size=128)
    sess = FusionSession(api, cfg, device="cuda")
    kfs = [scene.keyframe(k) for k in range(n)]
    for kf in kfs:
        sess.step(kf)
    return sess, cfg, kfs


def test_result_formats_match_reference_layout(tmp_path):
    sess, cfg, _ = _session()
    corners = port.GeneralInstance3DBoxes(sess.all_pred_box.pred_boxes_3d.tensor.cpu(), sess.all_pred_box.pred_boxes_3d.R.cpu()).corners.numpy()
    # tools/utils.py:302-317 restated: extents >= 0.3 on every axis
    ext = corners.max(1) - corners.min(1)
    expect = corners[(ext >= 0.3).all(1)]
    assert np.array_equal(results.post_process(corners), expect)
    assert np.array_equal(results.post_process(torch.from_numpy(corners).cuda()).cpu().numpy(), expect)
    assert 0 < len(expect) < len(corners)
    # demo.py:373-380 (scannet): [[(0, corners, 1.0)]] pickled with HIGHEST_PROTOCOL
    data = results.global_save_list(sess.all_pred_box, dataset="scannet")
    assert len(data) == 1 and len(data[0]) == len(expect)
    assert all(c == 0 and f == 1.0 and np.array_equal(b, e) for (c, b, f), e in zip(data[0], expect))
    path = tmp_path / "scene_boxes.pkl"
    results.save_box(data, path)
    assert open(path, "rb").read() == pickle.dumps([[(int(0), expect[n], 1.0) for n in range(len(expect))]], protocol=pickle.HIGHEST_PROTOCOL)
    back = results.load_data(path)
    assert np.array_equal(back[0][3][1], expect[3])
    # demo.py:383-387: framewise list carries class index and feature per observation
    m = len(sess.per_frame_ins)
    fw = results.framewise_save_list(sess.per_frame_ins, list(range(m)), [np.full(4, i, np.float32) for i in range(m)])
    pc = port.GeneralInstance3DBoxes(sess.per_frame_ins.pred_boxes_3d.tensor.cpu(), sess.per_frame_ins.pred_boxes_3d.R.cpu()).corners.numpy()
    assert len(fw[0]) == m and fw[0][2][0] == 2 and np.array_equal(fw[0][2][1], pc[2]) and fw[0][2][2][0] == 2.0


def test_record_and_replay(tmp_path):
    sess, cfg, kfs = _session(6)
    rec = replay.KeyframeRecorder()
    for i, kf in enumerate(kfs):
        if i % 2:
            rec.add_keyframe(kf)
        else:       # the torch-tensor path a detector-side hook would use
            rec.add(kf.frame_id, torch.from_numpy(kf.pose), kf.K, kf.image_size, torch.from_numpy(kf.tensor_cam).cuda(),
                    torch.from_numpy(kf.R_cam), kf.scores, kf.pred_boxes, kf.pred_proj_xy)
    path = str(tmp_path / "seq.npz")
    rec.save(path)
    frames = replay.load_sequence(path)
    assert len(frames) == 6 and all(np.array_equal(a.tensor_cam, b.tensor_cam) and np.array_equal(a.pose, b.pose) for a, b in zip(frames, kfs))
    # player: the recorded keyframes through the engine reproduce the live session
    eng = FusionEngine(cfg, map_capacity=256, store_capacity=1024, fused_capacity=256)
    for kf in frames:
        eng.step(pack_keyframe(kf.tensor_cam, kf.R_cam, kf.scores, kf.pred_boxes, kf.pred_proj_xy, kf.pose), kf.tensor_cam.shape[0], kf.K, kf.image_size)
    a, b = eng.snapshot(), sess.snapshot()
    for key in ("tensor", "scores", "fusion_flat", "fusion_off", "already_flat"):
        assert np.array_equal(a[key], b[key]), key


def test_detection_filter_matches_reference_masks():
    """the fused pre-filter equals demo.py:138-148 applied with the reference's torch code (the port)."""
    rs = np.random.RandomState(0)
    n = 4000
    dims = np.exp(rs.normal(np.log(0.4), 1.0, (n, 3))).astype(np.float32)
    dims[:200, 0] = 3.0; dims[:200, 1] = 0.05                                  # floor-like slabs
    dims[200:300] = np.array([0.14, 0.13, 1.2], np.float32) * rs.uniform(0.9, 1.1, (100, 3)).astype(np.float32)
    t = np.concatenate([rs.normal(0, 2, (n, 3)).astype(np.float32), dims], 1)
    uv = np.stack([rs.uniform(-20, 404, n), rs.uniform(-20, 532, n)], 1).astype(np.float32)
    sc = rs.uniform(0, 1, n).astype(np.float32)
    for shape, size_max in (("ca1m", 0), ("scannet", 2.5)):
        cfg = make_cfg(shape)
        cfg["detection"]["size_max_thres"] = size_max
        W, H = cfg["cam"]["W"], cfg["cam"]["H"]
        thr = cfg["detection"]["score_thresh"]
        ref_bm = port.BoxManager(cfg)
        tt, uu, ss = torch.from_numpy(t), torch.from_numpy(uv), torch.from_numpy(sc)
        expect = (ss >= float(thr)) & ref_bm.check_uv_bounds(uu, W, H, ratio=cfg["detection"]["uv_bound_value"]) & \
                 ~ref_bm.check_floor_mask(tt, ratio=cfg["detection"]["floor_ratio"])
        if size_max:
            expect = expect & ~ref_bm.check_large_mask(tt, thres=size_max)
        ins = api.Instances3D((H, W))
        ins.pred_boxes_3d = api.GeneralInstance3DBoxes(tt.cuda(), torch.eye(3).repeat(n, 1, 1).cuda())
        ins.pred_proj_xy, ins.scores = uu.cuda(), ss.cuda()
        got = api.BoxManager(cfg).filter_detections(ins, W, H, score_thresh=thr)
        assert torch.equal(got.cpu(), expect) and 0.05 < float(expect.float().mean()) < 0.95
