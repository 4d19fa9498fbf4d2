"""GPU tests of the device-resident engine: after every keyframe its state must equal what
the reference-shaped API - and the reference itself (goldens) - produce."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from boxfusion_b200 import api                                        # noqa: E402
from boxfusion_b200.driver import FusionSession                       # noqa: E402
from boxfusion_b200.engine import FusionEngine, pack_keyframe         # noqa: E402
from boxfusion_b200.synthetic import SyntheticScene, make_cfg, make_pst   # noqa: E402
from tests.golden.make_golden import SEQUENCES                        # noqa: E402

KEYS = ("tensor", "R", "scores", "valid_num", "init_id", "fusion_flat", "fusion_off", "fusion_flag", "already_flat", "already_off")


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8) if a.dtype.kind == "f" else a


def _pack(kf):
    return pack_keyframe(kf.tensor_cam, kf.R_cam, kf.scores, kf.pred_boxes, kf.pred_proj_xy, kf.pose)


@pytest.mark.parametrize("name", list(SEQUENCES))
def test_engine_matches_reference_golden(golden_dir, name):
    spec = dict(SEQUENCES[name])
    n_frames = spec.pop("frames")
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    scene = SyntheticScene(**spec)
    cfg = make_cfg(spec["shape"], pst_path=os.path.join(golden_dir, "pst_1024_0.npy"), pst_size=1024)
    eng = FusionEngine(cfg, map_capacity=512, store_capacity=4096, fused_capacity=1024)
    M = 0
    for k in range(n_frames):
        kf = scene.keyframe(k)
        n = kf.tensor_cam.shape[0]
        eng.step(_pack(kf), n, kf.K, kf.image_size)
        # the engine lifts and projects by itself: the stored observations must be the reference's tensors
        assert np.array_equal(_bits(eng.store["tensor"][M:M + n].cpu().numpy()), _bits(g[f"k{k}_tensor_w"]))
        np.testing.assert_allclose(eng.store["uv"][M:M + n].cpu().numpy().reshape(n, 8, 2), g[f"k{k}_projected"], rtol=1e-6, atol=1e-4)
        M += n
        snap = eng.snapshot()
        for key in KEYS:
            ref = g[f"k{k}_snap_{key}"]
            assert snap[key].shape == ref.shape and np.array_equal(_bits(snap[key]), _bits(ref)), (name, k, key)


def test_engine_matches_api_longer_sequence():
    """60 dense keyframes: the engine against the reference-shaped CUDA API (itself pinned to the reference)."""
    scene = SyntheticScene(n_objects=150, seed=9, max_det=45, shape="ca1m", tilt_noise=0.01)
    cfg = make_cfg("ca1m", pst_path=make_pst(512, seed=0), pst_size=512)
    eng = FusionEngine(cfg, map_capacity=2048, store_capacity=8192)
    sess = FusionSession(api, cfg, device="cuda")
    for k in range(60):
        kf = scene.keyframe(k)
        eng.step(_pack(kf), kf.tensor_cam.shape[0], kf.K, kf.image_size)
        sess.step(kf)
        a, b = eng.snapshot(), sess.snapshot()
        for key in KEYS:
            assert a[key].shape == b[key].shape and np.array_equal(_bits(a[key]), _bits(b[key])), (k, key)
    assert len(sess.box_manager.already_fusion) > 50 and eng.N == len(sess.all_pred_box)
    # export(): the reference-shaped containers rebuilt from the device state
    allp, per, bm = eng.export()
    assert bm.fusion_list == [[int(x) for x in l] for l in sess.box_manager.fusion_list]
    assert bm.already_fusion == [[int(x) for x in l] for l in sess.box_manager.already_fusion]
    assert torch.equal(allp.pred_boxes_3d.tensor.cpu(), sess.all_pred_box.pred_boxes_3d.tensor.cpu())
    assert torch.equal(per.projected_boxes.cpu(), sess.per_frame_ins.projected_boxes.cpu())


def test_check_valid_num_three_way():
    """cfg box_fusion.check_valid = True (BoxManager.check_valid_num, box_manager.py:151-166; off in the shipped configs):
    map boxes never re-observed are dropped after `gap` keyframes.  CPU port == reference-shaped CUDA API == engine."""
    from oracle import port
    port.IOU_BACKEND = "c"
    scene = SyntheticScene(n_objects=60, seed=21, max_det=20, shape="scannet", new_frac=0.35)
    cfg = make_cfg("scannet", pst_path=make_pst(256, seed=0), pst_size=256)
    cfg["box_fusion"]["check_valid"] = True
    cfg["data"]["gap"] = 2
    eng = FusionEngine(cfg, map_capacity=1024, store_capacity=4096)
    sess, ref = FusionSession(api, cfg, device="cuda"), FusionSession(port, cfg)
    dropped = 0
    for k in range(24):
        kf = scene.keyframe(k)
        eng.step(_pack(kf), kf.tensor_cam.shape[0], kf.K, kf.image_size)
        ins_b, pose_np = ref.make_pred_instances(kf)
        ins_a, _ = sess.pred_instances_from_world(kf, ins_b.pred_boxes_3d.tensor.numpy(), ins_b.pred_boxes_3d.R.numpy(),
                                                  ins_b.projected_boxes.numpy())
        before = len(sess.all_pred_box) if sess.all_pred_box is not None else 0
        sess.step(kf, ins_a, pose_np)
        ref.step(kf, ins_b, pose_np)
        a, b, c = eng.snapshot(), sess.snapshot(), ref.snapshot()
        for key in KEYS:
            assert b[key].shape == c[key].shape and np.array_equal(_bits(b[key]), _bits(c[key])), ("api vs port", k, key)
            assert a[key].shape == b[key].shape and np.array_equal(_bits(a[key]), _bits(b[key])), ("engine vs api", k, key)
        if sess.last_keep_idx is not None:
            dropped += max(0, len(sess.last_keep_idx) - len(sess.all_pred_box))
    assert dropped > 0, "the scene must exercise check_valid_num"


def test_frame_gap_sequence_matches_reference_golden(golden_dir):
    """A keyframe every 3rd frame with check_valid_num on (data.gap is in frames): the engine, given the
    frame index of every keyframe, and the reference-shaped API both equal the reference golden after every keyframe."""
    from tests.golden.make_golden import GAP_SEQUENCE
    name, spec, gap = GAP_SEQUENCE
    spec = dict(spec)
    n_frames = spec.pop("frames")
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = make_cfg(spec["shape"], pst_path=os.path.join(golden_dir, "pst_1024_0.npy"), pst_size=1024)
    cfg["data"]["gap"] = gap
    cfg["box_fusion"]["check_valid"] = True
    scene = SyntheticScene(**spec)
    eng = FusionEngine(cfg, map_capacity=512, store_capacity=4096, fused_capacity=1024)
    sess = FusionSession(api, cfg, device="cuda", frame_stride=gap)
    for k in range(n_frames):
        kf = scene.keyframe(k)
        eng.step(_pack(kf), kf.tensor_cam.shape[0], kf.K, kf.image_size, frame_id=gap * k)
        sess.step(kf)
        a, b = eng.snapshot(), sess.snapshot()
        for key in KEYS:
            ref = g[f"k{k}_snap_{key}"]
            assert a[key].shape == ref.shape and np.array_equal(_bits(a[key]), _bits(ref)), ("engine", k, key)
            assert b[key].shape == ref.shape and np.array_equal(_bits(b[key]), _bits(ref)), ("api", k, key)


def test_engine_graph_equals_eager_and_phase_by_phase():
    """The captured whole-keyframe graph, the same launches issued eagerly, and the keyframe issued phase by phase (what the
    reference-shaped API does) leave identical state."""
    from boxfusion_b200 import _lib
    scene = SyntheticScene(n_objects=80, seed=31, max_det=30, shape="ca1m", tilt_noise=0.01)
    cfg = make_cfg("ca1m", pst_path=make_pst(256, seed=2), pst_size=256)
    mk = lambda **kw: FusionEngine(cfg, map_capacity=512, store_capacity=2048, fused_capacity=512, **kw)   # noqa: E731
    a, b, c = mk(), mk(use_graph=False), mk()
    phases = [_lib.PH_INGEST, _lib.PH_NMS, _lib.PH_CORR, _lib.PH_COMPACT, _lib.PH_FUSE, _lib.PH_FINISH]
    for k in range(25):
        kf = scene.keyframe(k)
        n = kf.tensor_cam.shape[0]
        a.step(_pack(kf), n, kf.K, kf.image_size)
        b.step(_pack(kf), n, kf.K, kf.image_size)
        packed = _pack(kf)
        for i, ph in enumerate(phases):
            if n:
                c.step(packed, n, kf.K if i == 0 else None, kf.image_size if i == 0 else None, k if i == 0 else None, phases=ph)
        if not n:
            c.step(packed, 0)
        sa, sb, sc = a.snapshot(), b.snapshot(), c.snapshot()
        for key in KEYS:
            assert np.array_equal(_bits(sa[key]), _bits(sb[key])), ("graph vs eager", k, key)
            assert np.array_equal(_bits(sa[key]), _bits(sc[key])), ("graph vs phases", k, key)
    assert a.N > 20 and a.state().refine_boxes_total > 5 and a.M == c.M == b.M


def test_engine_reset_reuses_buffers_and_graphs():
    scene = SyntheticScene(n_objects=40, seed=3, max_det=16)
    cfg = make_cfg("ca1m", pst_path=make_pst(256, seed=0), pst_size=256)
    eng = FusionEngine(cfg, map_capacity=256, store_capacity=1024, fused_capacity=256)
    snaps = []
    for rep in range(2):
        for k in range(8):
            kf = scene.keyframe(k)
            eng.step(_pack(kf), kf.tensor_cam.shape[0], kf.K, kf.image_size)
        snaps.append(eng.snapshot())
        eng.reset()
        assert eng.N == 0 and eng.M == 0
    for key in KEYS:
        assert np.array_equal(_bits(snaps[0][key]), _bits(snaps[1][key])), key


def test_engine_empty_keyframe_and_capacity():
    scene = SyntheticScene(n_objects=20, seed=2, max_det=8)
    cfg = make_cfg("ca1m", pst_path=make_pst(64), pst_size=64)
    eng = FusionEngine(cfg, map_capacity=16, store_capacity=64, fused_capacity=16)
    kf = scene.keyframe(0)
    eng.step(_pack(kf), kf.tensor_cam.shape[0], kf.K, kf.image_size)
    n0 = eng.N
    eng.step(np.zeros(48, np.float32), 0, kf.K, kf.image_size)           # empty keyframe: demo.py:206-212
    assert eng.N == n0 and eng.count == 2
    with pytest.raises(RuntimeError):
        for k in range(1, 20):
            kf = scene.keyframe(k)
            eng.step(_pack(kf), kf.tensor_cam.shape[0], kf.K, kf.image_size)


def test_concurrent_engines_on_private_streams():
    """Three independent sequences interleaved from one host thread (step_launch / step_finish on private streams and
    handles) end in exactly the state they reach when run one after another."""
    cfg = make_cfg("scannet", pst_path=make_pst(256, seed=3), pst_size=256)
    scenes = [SyntheticScene(n_objects=60, seed=20 + i, max_det=25, shape="scannet") for i in range(3)]
    solo = []
    for sc in scenes:
        e = FusionEngine(cfg, map_capacity=512, store_capacity=2048, fused_capacity=512)
        for k in range(15):
            kf = sc.keyframe(k)
            e.step(_pack(kf), kf.tensor_cam.shape[0], kf.K, kf.image_size)
        solo.append(e.snapshot())
    engines = [FusionEngine(cfg, map_capacity=512, store_capacity=2048, fused_capacity=512, private_stream=True) for _ in scenes]
    for k in range(15):
        kfs = [sc.keyframe(k) for sc in scenes]
        for e, kf in zip(engines, kfs):
            e.step_launch(_pack(kf), kf.tensor_cam.shape[0], kf.K, kf.image_size)
        for e in engines:
            e.step_finish()
    for e, ref in zip(engines, solo):
        got = e.snapshot()
        for key in KEYS:
            assert np.array_equal(_bits(got[key]), _bits(ref[key])), key
