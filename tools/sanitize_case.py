"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): a short sequence through the reference-shaped
API and the device-resident engine, an IoU matrix in both modes, and a refinement launch with a multi-CTA cluster."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from boxfusion_b200 import api, ops                                     # noqa: E402
from boxfusion_b200.driver import FusionSession                         # noqa: E402
from boxfusion_b200.engine import FusionEngine, pack_keyframe           # noqa: E402
from boxfusion_b200.synthetic import SyntheticScene, make_cfg, make_pst, map_and_detections   # noqa: E402

scene = SyntheticScene(n_objects=40, seed=3, max_det=16, tilt_noise=0.01)
cfg = make_cfg("ca1m", pst_path=make_pst(256, seed=0), pst_size=256)
sess = FusionSession(api, cfg, device="cuda")
eng = FusionEngine(cfg, map_capacity=256, store_capacity=1024, fused_capacity=256)
for k in range(8):
    kf = scene.keyframe(k)
    sess.step(kf)
    eng.step(pack_keyframe(kf.tensor_cam, kf.R_cam, kf.scores, kf.pred_boxes, kf.pred_proj_xy, kf.pose), kf.tensor_cam.shape[0], kf.K, kf.image_size)
a, b = eng.snapshot(), sess.snapshot()
assert all(np.array_equal(a[k], b[k]) for k in ("tensor", "fusion_flat", "already_flat"))
(mt, mR, _), (dt, dR, _) = map_and_detections(300, 60, seed=2, tilt_noise=0.0)
ca, cb = ops.box_corners(dt, dR), ops.box_corners(mt, mR)
for mode in (ops.IOU_SAMPLED_REF, ops.IOU_ANALYTIC):
    ops.iou3d_matrix(ca, cb, mode=mode, want_counts=True, want_stats=True)
# the out-of-line redo of the branch-free divisions (bf_eval_view_cold): camera-frame x of four corners is 1e-30
bf = api.BoxFusion(cfg)
bf.update_intrinsics((384, 512), np.array([[500.0, 0, 192.0], [0, 500.0, 256.0], [0, 0, 1]], np.float32))
box6 = np.array([0.5, 0.1, 0.2, 1.0, 0.6, 0.8], np.float32)
poses = np.tile(np.eye(4, dtype=np.float32), (2, 1, 1))
poses[:, :3, 3] = (-1e-30, 0.0, -3.0)
uv = np.tile(np.array([[192, 224], [352, 224], [352, 320], [192, 320], [192, 231], [315, 231], [315, 305], [192, 305]], np.float32), (2, 1, 1))
ops.cold_redos()
bf.evaluate_iou(box6.astype(np.float64), uv, np.eye(3, dtype=np.float32), np.ones(2, np.float32), poses,
                np.array([0.0, 0.1, 0.1, 0.0, 0.5, 0.5], np.float32), 2)
assert ops.cold_redos() == 2 * 256
torch.cuda.synchronize()
print("sanitize case ok: map", eng.N, "fused", len(sess.box_manager.already_fusion))
