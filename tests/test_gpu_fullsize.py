"""GPU parity at the FULL sizes of BASELINE.json's configs (round 2): the exact sequence bench.py times,
the 4 352-box NMS of configs[2] and every one of the 128 boxes of configs[3], against the CPU oracle (oracle/port.py with its C
backend, OpenMP over pairs / particles - bit-identical to the scipy/Qhull backend that is pinned to the reference)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bench                                                            # noqa: E402
from boxfusion_b200 import api, ops                                     # noqa: E402
from boxfusion_b200.driver import FusionSession                         # noqa: E402
from boxfusion_b200.engine import FusionEngine, pack_keyframe           # noqa: E402
from boxfusion_b200.synthetic import make_cfg, make_pst, map_and_detections, refine_problem   # noqa: E402
from oracle import port, refine_oracle as ro                            # noqa: E402

KEYS = ("tensor", "R", "scores", "valid_num", "init_id", "fusion_flat", "fusion_off", "fusion_flag", "already_flat", "already_off")


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8) if a.dtype.kind == "f" else a


def test_bench_sequence_api_and_engine_match_port():
    """The 300-keyframe CA-1M-shaped sequence of bench.py (rank 0: seed 1, 200 objects, <= 50 detections, shipped template)
    through the reference-shaped API (engine-backed fast path) and through FusionEngine, field by field against the port
    after EVERY keyframe, incl. the keep indices the API returns."""
    port.IOU_BACKEND = "c_batch"
    ro.set_threads(os.cpu_count() or 1)
    cfg = make_cfg("ca1m", pst_path=bench.GOLDEN_PST, pst_size=1024)
    frames = bench.build_keyframes(1)
    a, b = FusionSession(api, cfg, device="cuda"), FusionSession(port, cfg)
    eng = FusionEngine(cfg, map_capacity=4096, store_capacity=65536)
    fused = 0
    for k, kf in enumerate(frames):
        ka, kb = a.step(kf), b.step(kf)
        eng.step(pack_keyframe(kf.tensor_cam, kf.R_cam, kf.scores, kf.pred_boxes, kf.pred_proj_xy, kf.pose, kf.K, kf.image_size, k),
                 kf.tensor_cam.shape[0])
        assert (ka is None and kb is None) or np.array_equal(ka, kb), k
        if k % 5 == 4 or k > 290:
            sa, sb, se = a.snapshot(), b.snapshot(), eng.snapshot()
            for key in KEYS:
                assert sa[key].shape == sb[key].shape and np.array_equal(_bits(sa[key]), _bits(sb[key])), ("api", k, key)
                assert se[key].shape == sb[key].shape and np.array_equal(_bits(se[key]), _bits(sb[key])), ("engine", k, key)
    fused = len(b.box_manager.already_fusion)
    assert len(b.all_pred_box) > 300 and fused > 400 and a.box_manager._session is not None


def test_c3_nms_4352_matches_port():
    """BASELINE configs[2]: 3-D NMS over the 4 096-box map + 256 detections - keep, success (valid_num), fusion lists and
    flags equal the port's, with multi-view lists on a quarter of the map so that record()'s merge / swap branches run."""
    port.IOU_BACKEND = "c_batch"
    ro.set_threads(os.cpu_count() or 1)
    (mt, mR, ms_), (dt, dR, ds) = map_and_detections(4096, 256, seed=3, tilt_noise=0.0)
    t = np.concatenate([mt, dt]); R = np.concatenate([mR, dR]); sc = np.concatenate([ms_, ds])
    n = t.shape[0]
    rs = np.random.RandomState(3)
    lists, M = [], 0
    for i in range(n):
        k = 1 if i >= 4096 else int(rs.choice([1, 1, 1, 2, 3]))
        lists.append(list(range(M, M + k))); M += k
    poses = np.tile(np.eye(4, dtype=np.float32), (M, 1, 1))
    ang = np.deg2rad(rs.uniform(0, 90, M))
    poses[:, 0, 0], poses[:, 0, 1], poses[:, 1, 0], poses[:, 1, 1] = np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)
    poses[:, :3, 3] = rs.uniform(-1.5, 1.5, (M, 3)).astype(np.float32)
    flags = [int(rs.rand() < 0.2) if len(l) > 1 else 0 for l in lists]
    init_id = np.array([l[0] for l in lists], dtype=np.int64)
    cfg = make_cfg("ca1m", pst_path=make_pst(32))
    out = []
    for impl, dev in ((api, "cuda"), (port, "cpu")):
        bm = impl.BoxManager(cfg)
        bm.fusion_list = [list(l) for l in lists]
        bm.fusion_flag = list(flags)
        ins = impl.Instances3D((512, 384))
        ins.pred_boxes_3d = impl.GeneralInstance3DBoxes(torch.from_numpy(t).to(dev), torch.from_numpy(R).to(dev))
        ins.scores = torch.from_numpy(sc).to(dev)
        ins.init_id = torch.from_numpy(init_id)
        ins.valid_num = torch.zeros(n, device=dev)
        keep, succ = impl.Instances3D.spatial_association(ins, 0.1, bm, torch.from_numpy(poses))
        out.append(([int(x) for x in keep], [int(x) for x in succ], [[int(x) for x in l] for l in bm.fusion_list], list(bm.fusion_flag),
                    ins.valid_num.cpu().numpy().copy()))
    got, ref = out
    assert got[0] == ref[0] and got[1] == ref[1] and len(ref[1]) > 300 and len(ref[0]) < n - 300
    assert got[2] == ref[2] and got[3] == ref[3] and np.array_equal(got[4], ref[4])
    assert max(len(l) for l in ref[2]) >= 4


def test_c4_every_box_matches_oracle():
    """BASELINE configs[3]: 4096 particles x 32 views x 128 boxes (the reference's early stop on) - updated flag, iteration
    count and fused box of ALL 128 boxes equal the C oracle's, bit for bit."""
    ro.set_threads(os.cpu_count() or 1)
    B, V, P = 128, 32, 4096
    prob = refine_problem(B, V, seed=11)
    W, H = prob["size"]
    pst = make_pst(P, seed=1)
    cfg = make_cfg("ca1m", pst_path=pst, pst_size=P)
    K16 = ro.K16_from_K3(prob["K"])
    corners = ops.box_corners(prob["tensor"].reshape(-1, 6), prob["R"].reshape(-1, 3, 3))
    proj = ops.project_boxes(corners, torch.linalg.inv(torch.from_numpy(prob["poses"].reshape(-1, 4, 4))), prob["K"], W, H)
    proj_h = proj.cpu().numpy().reshape(B, V, 16)
    rcfg = ops.make_refine_cfg(cfg, K16, H, W)
    off = np.arange(B + 1, dtype=np.int32) * V
    idx = np.arange(B * V, dtype=np.int32)
    out, upd, its, _, status = ops.refine(pst, prob["tensor"].reshape(-1, 6), prob["R"].reshape(-1, 9), prob["scores"].reshape(-1), proj,
                                          prob["poses"].reshape(-1, 16), off, idx, rcfg, max_views=V)
    assert int(status.item()) == 0 and ops.last_refine_launch()["variant"] == "saturated"
    out, upd, its = out.cpu().numpy(), upd.cpu().numpy(), its.cpu().numpy()
    cs = ro.make_cfg_struct(cfg, H, W)
    for b in range(B):
        u, o6, n_it, _ = ro.refine_box(prob["tensor"][b], prob["R"][b], prob["scores"][b], proj_h[b], prob["poses"][b], pst, K16, cs,
                                       want_trace=True)
        assert bool(upd[b]) == u and int(its[b]) == n_it, b
        if u:
            assert np.array_equal(_bits(out[b]), _bits(o6)), b
