"""GPU tests of the reference-shaped API running on the device-resident engine (boxfusion_b200/fastpath.py): the calls
demo.py:243-327 makes, with containers that are lazy views of the engine state, must leave exactly the state the
call-by-call implementation - and the CPU port of the reference - leave, also when the caller strays from demo.py's
pattern half-way through a keyframe."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from boxfusion_b200 import api, fastpath                              # noqa: E402
from boxfusion_b200.driver import FusionSession                       # noqa: E402
from boxfusion_b200.synthetic import SyntheticScene, make_cfg, make_pst   # noqa: E402
from oracle import port                                               # noqa: E402


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8) if a.dtype.kind == "f" else a


def _same(sa, sb, what):
    for key in sa:
        assert sa[key].shape == sb[key].shape and np.array_equal(_bits(sa[key]), _bits(sb[key])), (what, key)


def _pair(scene_kw, cfg, frames, hook=None, extras=False):
    """The CUDA API (fast path allowed) against the CPU port, every mutated field after every keyframe."""
    port.IOU_BACKEND = "c"
    scene = SyntheticScene(**scene_kw)
    a, b = FusionSession(api, cfg, device="cuda"), FusionSession(port, cfg)
    fast_frames = 0
    for k in range(frames):
        kf = scene.keyframe(k)
        if kf.tensor_cam.shape[0] == 0:
            a.step(kf); b.step(kf)
            continue
        ins_b, pose_np = b.make_pred_instances(kf)
        ins_a, _ = a.make_pred_instances(kf)
        if extras:
            n = len(ins_a)
            cats = np.array([f"c{k}_{i}" for i in range(n)])
            feat = torch.arange(n * 4, dtype=torch.float32).reshape(n, 4) + 100 * k
            ins_a.categories, ins_b.categories = cats.copy(), cats.copy()
            ins_a.features, ins_b.features = feat.cuda(), feat.clone()
        if hook is not None:
            hook(k, a)
        was_fast = a.box_manager._session is not None
        a.step(kf, ins_a, pose_np); b.step(kf, ins_b, pose_np)
        fast_frames += int(was_fast and a.box_manager._session is not None)
        _same(a.snapshot(), b.snapshot(), k)
        if extras:
            assert list(a.all_pred_box.categories) == list(b.all_pred_box.categories), k
            assert torch.equal(a.all_pred_box.features.cpu(), b.all_pred_box.features), k
            assert list(a.per_frame_ins.categories) == list(b.per_frame_ins.categories), k
    return a, b, fast_frames


def test_fast_path_matches_port_and_is_taken():
    cfg = make_cfg("ca1m", pst_path=make_pst(512, seed=0), pst_size=512)
    a, b, fast = _pair(dict(n_objects=120, seed=7, max_det=40, shape="ca1m", tilt_noise=0.01), cfg, 30)
    assert fast >= 25, "the engine-backed path must carry the sequence"
    assert len(b.box_manager.already_fusion) > 20
    # the containers are views of the engine state
    assert isinstance(a.all_pred_box._fields, fastpath.EngineFields) and a.all_pred_box.pred_boxes_3d.tensor.is_cuda
    assert len(a.all_pred_box) == len(b.all_pred_box) and len(a.per_frame_ins) == len(b.per_frame_ins)
    assert torch.equal(a.per_frame_ins.projected_boxes.cpu(), b.per_frame_ins.projected_boxes)
    assert torch.equal(a.per_frame_ins.init_id.cpu(), b.per_frame_ins.init_id)


def test_fast_path_equals_call_by_call_path(monkeypatch):
    """Same sequence with the fast path switched off: identical state and identical returned indices."""
    cfg = make_cfg("scannet", pst_path=make_pst(256, seed=4), pst_size=256)
    scene = SyntheticScene(n_objects=90, seed=12, max_det=30, shape="scannet", tilt_noise=0.01)
    fastpath.ENABLED = True
    a = FusionSession(api, cfg, device="cuda")
    keep_a, snap_a = [], []
    for k in range(25):
        keep_a.append(a.step(scene.keyframe(k)))
        snap_a.append(a.snapshot())
    assert a.box_manager._session is not None
    monkeypatch.setattr(fastpath, "ENABLED", False)
    b = FusionSession(api, cfg, device="cuda")
    for k in range(25):
        kb = b.step(scene.keyframe(k))
        assert (kb is None and keep_a[k] is None) or np.array_equal(kb, keep_a[k]), k
        _same(snap_a[k], b.snapshot(), k)
    assert b.box_manager._session is None


def test_fast_path_with_extra_fields():
    """Fields the engine does not hold (categories: numpy strings, features: tensors - demo.py:168-170) ride along."""
    cfg = make_cfg("ca1m", pst_path=make_pst(256, seed=1), pst_size=256)
    _, _, fast = _pair(dict(n_objects=60, seed=3, max_det=25, shape="ca1m"), cfg, 16, extras=True)
    assert fast >= 10


def test_leaving_the_fast_path_mid_sequence():
    """A caller that strays from demo.py's pattern: edits a fusion list, replaces a list, mutates the map in place, reads
    everything.  The state is exported, the call-by-call path carries on, and the next boxfusion() re-enters."""
    cfg = make_cfg("ca1m", pst_path=make_pst(256, seed=2), pst_size=256)
    port.IOU_BACKEND = "c"
    scene = SyntheticScene(n_objects=80, seed=5, max_det=30, shape="ca1m", tilt_noise=0.01)
    a, b = FusionSession(api, cfg, device="cuda"), FusionSession(port, cfg)
    left = 0
    for k in range(24):
        kf = scene.keyframe(k)
        if k in (8, 15):                                     # edit one list on both sides (the reference's lists are plain Python)
            for s in (a, b):
                fl = s.box_manager.fusion_list
                i = max(range(len(fl)), key=lambda j: len(fl[j]))
                fl[i].append(fl[i][-1])                      # a duplicate observation: changes what gets fused later
            left += 1
        if k == 11:                                          # wholesale replacement
            for s in (a, b):
                s.box_manager.fusion_flag = list(s.box_manager.fusion_flag)
        if k == 18:                                          # in-place edit of the map through the handed-out view
            for s in (a, b):
                s.all_pred_box.pred_boxes_3d.tensor[0, 3:6] *= 1.25
                s.all_pred_box.scores[1] += 0.01
        ins_b, pose_np = b.make_pred_instances(kf)
        ins_a, _ = a.make_pred_instances(kf)
        a.step(kf, ins_a, pose_np); b.step(kf, ins_b, pose_np)
        _same(a.snapshot(), b.snapshot(), k)
    assert left == 2 and a.box_manager._session is not None   # re-entered after every exit


def test_fast_path_other_threshold_falls_back():
    """spatial_association with a threshold other than cfg's cannot use the engine's captured keyframe: call-by-call result."""
    cfg = make_cfg("ca1m", pst_path=make_pst(256, seed=2), pst_size=256)
    port.IOU_BACKEND = "c"
    scene = SyntheticScene(n_objects=50, seed=9, max_det=20, shape="ca1m")
    cfg2 = make_cfg("ca1m", pst_path=make_pst(256, seed=2), pst_size=256)
    a, b = FusionSession(api, cfg, device="cuda"), FusionSession(port, cfg2)
    for k in range(14):
        kf = scene.keyframe(k)
        if k == 9:                                           # the driver reads the threshold from cfg at call time
            cfg["box_fusion"]["nms_threshold"] = 0.25
            cfg2["box_fusion"]["nms_threshold"] = 0.25
        ins_b, pose_np = b.make_pred_instances(kf)
        ins_a, _ = a.make_pred_instances(kf)
        a.step(kf, ins_a, pose_np); b.step(kf, ins_b, pose_np)
        _same(a.snapshot(), b.snapshot(), k)


def test_run_ahead_is_invisible():
    """spatial_association queues the rest of the keyframe behind itself (fastpath.RUN_AHEAD).  A caller that looks at the
    state between its calls, or stops following demo.py half-way through a keyframe, must see exactly what it would have seen
    without it: the engine is rolled back to where the caller is."""
    cfg = make_cfg("ca1m", pst_path=make_pst(256, seed=5), pst_size=256)
    port.IOU_BACKEND = "c"
    scene = SyntheticScene(n_objects=70, seed=14, max_det=28, shape="ca1m", tilt_noise=0.01)
    a, b = FusionSession(api, cfg, device="cuda"), FusionSession(port, cfg)
    looked = 0

    def step_with_peeks(sess, kf, peek):
        """driver._step with observations between the calls (what demo.py:294 does with its print of fusion_list)."""
        impl, bm = sess.impl, sess.box_manager
        ins, pose_np = sess.make_pred_instances(kf)
        count = sess.count
        sess.box_fuser.update_intrinsics(kf.image_size, kf.K)
        sess.all_kf_pose[count] = kf.pose
        sess.box_count += len(ins)
        bm.init_new_predictions(len(ins), len(sess.per_frame_ins))
        nb = len(sess.all_pred_box)
        cur_global = sess.all_pred_box
        allp = impl.Instances3D.cat([sess.all_pred_box, ins])
        sess.per_frame_ins = impl.Instances3D.cat([sess.per_frame_ins, ins])
        all_poses = np.concatenate((sess.all_poses, pose_np), axis=0)
        mask, succ = impl.Instances3D.spatial_association(allp, cfg["box_fusion"]["nms_threshold"], bm, sess.per_frame_ins.cam_pose)
        seen = {}
        if peek == "after_nms":                              # lists and valid_num right after nms_3d
            seen["fl"] = [list(l) for l in bm.fusion_list]
            seen["valid"] = allp.valid_num.cpu().numpy().copy()
        ck = [i - nb for i in mask if i >= nb]
        cs = [i - nb for i in succ if i >= nb]
        keep_idx = np.asarray(mask)
        if len(ck) > 0:
            allp, all_poses, keep_idx = impl.Instances3D.correspondence_association(
                cfg, bm, ck, cs, ins, cur_global, allp, all_poses, sess.per_frame_ins.cam_pose, count, mask, torch.from_numpy(kf.K),
                sess.all_kf_pose, threshold=cfg["association"]["small_threshold"], H=kf.image_size[1], W=kf.image_size[0])
            bm.update(keep_idx)
            if peek == "after_update":                       # demo.py:294; and the map before it is fused
                seen["fl"] = [list(l) for l in bm.fusion_list]
                seen["tensor"] = allp.pred_boxes_3d.tensor.cpu().numpy().copy()
                seen["flag"] = list(bm.fusion_flag)
            if peek != "skip_fusion":
                sess.box_fuser.boxfusion(allp, sess.per_frame_ins, bm)
        else:
            allp = allp[mask]
            all_poses = all_poses[mask]
            bm.update(keep_idx)
        sess.all_pred_box, sess.all_poses = allp, all_poses
        sess.count += 1
        return seen

    for k in range(22):
        kf = scene.keyframe(k)
        peek = {6: "after_nms", 9: "after_update", 12: "skip_fusion", 15: "after_update", 16: "after_nms"}.get(k)
        if k < 2 or peek is None:
            a.step(kf); b.step(kf)
        else:
            fast_before = a.box_manager._session is not None
            sa, sb = step_with_peeks(a, kf, peek), step_with_peeks(b, kf, peek)
            looked += int(fast_before)
            assert sa.keys() == sb.keys()
            for key in sa:
                if key in ("fl", "flag"):
                    assert sa[key] == [[int(x) for x in l] for l in sb[key]] if key == "fl" else sa[key] == [int(x) for x in sb[key]], (k, key)
                else:
                    assert np.array_equal(_bits(sa[key]), _bits(np.asarray(sb[key]))), (k, key)
        _same(a.snapshot(), b.snapshot(), k)
    assert looked >= 4 and a.box_manager._session is not None


def test_fast_path_without_graphs(monkeypatch):
    """The engine behind the API with use_graph=False: every phase and the run-ahead (NMS + flag publication, snapshot +
    correspondence + flag publication, the rest) issued as plain launches instead of captured graphs - same state."""
    monkeypatch.setattr(fastpath, "ENGINE_KWARGS", {"use_graph": False})
    cfg = make_cfg("ca1m", pst_path=make_pst(256, seed=6), pst_size=256)
    a, _, fast = _pair(dict(n_objects=60, seed=21, max_det=24, shape="ca1m", tilt_noise=0.01), cfg, 14)
    assert fast >= 9 and a.box_manager._session is not None
