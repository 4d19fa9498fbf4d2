"""GPU tests at BASELINE.json's full sizes through size-independent properties, plus edge cases the
reference's data can produce (empty keyframes, a single box, everything suppressed, degenerate boxes)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from boxfusion_b200 import api, ops                                   # noqa: E402
from boxfusion_b200.driver import FusionSession                       # noqa: E402
from boxfusion_b200.synthetic import (Keyframe, SyntheticScene, make_cfg, make_pst, map_and_detections,   # noqa: E402
                                      random_boxes, refine_problem)
from oracle import port, refine_oracle as ro                          # noqa: E402


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8) if a.dtype.kind == "f" else a


# ---- C3: 256 x 4096 IoU matrix + NMS over N = 4352 --------------------------------------------------------

@pytest.fixture(scope="module")
def c3():
    (mt, mR, ms), (dt, dR, ds) = map_and_detections(4096, 256, seed=3, tilt_noise=0.01)
    return ops.box_corners(dt, dR), ops.box_corners(mt, mR), (mt, mR, ms), (dt, dR, ds)


def test_c3_iou_matrix_properties(c3):
    ca, cb, _, _ = c3
    iou, cnt, stats = ops.iou3d_matrix(ca, cb, want_counts=True, want_stats=True)
    assert iou.shape == (256, 4096)
    assert float(iou.min()) >= 0.0 and float(iou.max()) <= 1.0
    # symmetry: swapping the operands swaps count1/count2 and leaves common and the IoU unchanged, bit for bit
    iou_t, cnt_t = ops.iou3d_matrix(cb, ca, want_counts=True)
    assert torch.equal(iou, iou_t.T)
    assert torch.equal(cnt[..., 2], cnt_t[..., 2].T) and torch.equal(cnt[..., 0], cnt_t[..., 1].T)
    # counts are consistent: common <= min(count1, count2) <= 25^3, zero rows exactly where the gate failed
    c = cnt.cpu().numpy().astype(np.int64)
    assert (c[..., 2] <= np.minimum(c[..., 0], c[..., 1])).all() and c.max() <= 25 ** 3
    assert int(stats[0]) == 256 * 4096 and int((c.sum(-1) > 0).sum()) == int(stats[2])
    # every gate-passing pair and 3000 random pairs against the CPU oracle, exactly
    ia, ib = np.nonzero(c.sum(-1) > 0)
    rs = np.random.RandomState(0)
    ia = np.concatenate([ia, rs.randint(0, 256, 3000)]); ib = np.concatenate([ib, rs.randint(0, 4096, 3000)])
    allc = np.concatenate([ca.cpu().numpy(), cb.cpu().numpy()])
    gate, ref = port.obb_counts_pairs_c(allc, ia, ib + 256)
    assert np.array_equal(c[ia, ib], ref.astype(np.int64))
    assert np.array_equal(iou.cpu().numpy()[ia, ib], np.where(gate > 0, port.iou_from_counts(ref), 0.0))


def test_c3_self_iou_is_one(c3):
    _, cb, _, _ = c3
    sub = cb[:512]
    d = torch.diagonal(ops.iou3d_matrix(sub, sub))
    assert float(d.min()) > 0.999                                      # c/(c+1e-6)


def test_c3_nms_kept_set_is_independent(c3):
    """With singleton fusion lists there are no keep swaps, so the kept set must contain no pair above the
    threshold, every dropped box must overlap a kept box with a higher score, and a second pass keeps all."""
    ca, cb, (mt, mR, ms), (dt, dR, ds) = c3
    corners = torch.cat([cb, ca]); n = corners.shape[0]
    scores = torch.from_numpy(np.concatenate([ms, ds])).cuda()
    cen = corners.mean(dim=1).contiguous()
    order = torch.argsort(scores, descending=True, stable=True).to(torch.int32)
    iid = torch.arange(n, dtype=torch.int32, device="cuda")
    poses = torch.eye(4, device="cuda").reshape(1, 16).repeat(n, 1).contiguous()

    def run(c, o, k):
        fl = torch.zeros((k, ops.FUSION_CAP), dtype=torch.int32, device="cuda"); fl[:, 0] = iid[:k]
        ln = torch.ones(k, dtype=torch.int32, device="cuda"); fg = torch.zeros(k, dtype=torch.int32, device="cuda")
        keep, succ, status = ops.nms3d(c, cen[:k], o, iid[:k], poses[:k], fl, ln, fg, 0.1, 0.8, 30.0, 0.5)
        assert int(status.item()) == 0
        return keep.bool(), succ.bool(), ln

    keep, succ, ln = run(corners, order, n)
    kept = torch.nonzero(keep).flatten()
    assert 0 < kept.numel() < n and int(succ.sum()) > 100
    kc = corners[kept]
    m = ops.iou3d_matrix(kc, kc)
    m.fill_diagonal_(0.0)
    assert float(m.max()) <= 0.1                                       # independence of the kept set
    dropped = torch.nonzero(~keep).flatten()
    cross = ops.iou3d_matrix(corners[dropped], kc)                     # every dropped box has a better-scoring kept partner
    better = scores[kept][None, :] > scores[dropped][:, None]
    assert bool(((cross > 0.1) & better).any(dim=1).all())
    # idempotence
    o2 = torch.argsort(scores[kept], descending=True, stable=True).to(torch.int32)
    keep2, succ2, _ = run(kc.contiguous(), o2, kept.numel())
    assert bool(keep2.all()) and int(succ2.sum()) == 0


# ---- C4: 4096 particles x 32 views x 128 boxes ------------------------------------------------------------

def test_c4_refine_properties():
    B, V, P = 128, 32, 4096
    prob = refine_problem(B, V, seed=11)
    W, H = prob["size"]
    pst = make_pst(P, seed=1)
    cfg = make_cfg("ca1m", pst_path=pst, pst_size=P)
    K16 = ro.K16_from_K3(prob["K"])
    t = torch.from_numpy(prob["tensor"].reshape(-1, 6)).cuda(); R = torch.from_numpy(prob["R"].reshape(-1, 9)).cuda()
    s = torch.from_numpy(prob["scores"].reshape(-1)).cuda(); po = torch.from_numpy(prob["poses"].reshape(-1, 16)).cuda()
    uv = ops.project_boxes(ops.box_corners(t, R), torch.linalg.inv(po.reshape(-1, 4, 4)), prob["K"], W, H).reshape(-1, 16)
    off = torch.arange(B + 1, dtype=torch.int32, device="cuda") * V
    idx = torch.arange(B * V, dtype=torch.int32, device="cuda")
    rcfg = ops.make_refine_cfg(cfg, K16, H, W)
    out, upd, its, trace, status = ops.refine(torch.from_numpy(pst).cuda(), t, R, s, uv, po, off, idx, rcfg, want_trace=True, max_views=V)
    assert int(status.item()) == 0
    out2, upd2, its2, trace2, _ = ops.refine(torch.from_numpy(pst).cuda(), t, R, s, uv, po, off, idx, rcfg, want_trace=True, max_views=V)
    assert torch.equal(out, out2) and torch.equal(its, its2) and torch.equal(trace, trace2)      # deterministic
    its_h, tr, out_h = its.cpu().numpy(), trace.cpu().numpy(), out.cpu().numpy()
    assert (its_h >= 3).all() and (its_h <= 20).all() and bool(upd.bool().any())
    for b in range(B):
        k = its_h[b]
        assert set(np.unique(tr[b, :k, 0])) <= {0.0, 1.0}
        assert (tr[b, :k, 1] >= 0).all() and (tr[b, :k, 1] <= 1.0 + 1e-6).all()              # mean |1-iou| in [0,1]
        assert (tr[b, :k, 2:] >= 1e-3).all()                                                   # radii never below min_scale
        if k < 20:
            assert (tr[b, k - 3:k, 0] == 0).all()                                              # stopped on 3 failures
    assert (out_h[upd.cpu().numpy() != 0, 3:] >= 0.01).all()                                  # dims floor (:719)
    # two boxes against the CPU oracle, bit for bit
    proj_h = uv.cpu().numpy().reshape(B, V, 16)
    cs = ro.make_cfg_struct(cfg, H, W)
    for b in (0, 77):
        u, o6, n_it, trr = ro.refine_box(prob["tensor"][b], prob["R"][b], prob["scores"][b], proj_h[b], prob["poses"][b],
                                         pst, K16, cs, want_trace=True)
        assert bool(upd[b]) == u and int(its_h[b]) == n_it
        assert np.array_equal(_bits(tr[b, :n_it]), _bits(trr[:n_it]))
        if u:
            assert np.array_equal(_bits(out_h[b]), _bits(o6))


# ---- edge cases ---------------------------------------------------------------------------------------------

def _empty_keyframe(kf):
    z = lambda *s: np.zeros(s, np.float32)
    return Keyframe(frame_id=kf.frame_id, pose=kf.pose, K=kf.K, image_size=kf.image_size, tensor_cam=z(0, 6), R_cam=z(0, 3, 3),
                    scores=z(0), pred_boxes=z(0, 4), pred_proj_xy=z(0, 2), gt_index=np.zeros(0, np.int64))


def test_sequence_with_empty_single_and_repeated_keyframes(monkeypatch):
    """Empty detections (demo.py:206-212), a single detection, and the same keyframe twice (every new box suppressed,
    the `no new box` branch demo.py:324-328) - against the CPU port."""
    monkeypatch.setattr(port, "IOU_BACKEND", "c")
    scene = SyntheticScene(n_objects=30, seed=4, max_det=12)
    cfg = make_cfg("ca1m", pst_path=make_pst(128, seed=2), pst_size=128)
    a, b = FusionSession(api, cfg, device="cuda"), FusionSession(port, cfg)
    again = scene.keyframe(2)
    again.scores = (again.scores * np.float32(0.97)).astype(np.float32)     # same boxes, lower scores: no exact score ties
    kfs = [scene.keyframe(0), _empty_keyframe(scene.keyframe(1)), scene.keyframe(2), again, scene.keyframe(3)]
    one = scene.keyframe(4)
    for f in ("tensor_cam", "R_cam", "scores", "pred_boxes", "pred_proj_xy", "gt_index"):
        setattr(one, f, getattr(one, f)[:1])
    kfs.append(one)
    for k, kf in enumerate(kfs):
        if kf.tensor_cam.shape[0] == 0:
            assert a.step(kf) is None and b.step(kf) is None
            continue
        ins_b, pose_np = b.make_pred_instances(kf)
        ins_a, _ = a.pred_instances_from_world(kf, ins_b.pred_boxes_3d.tensor.numpy(), ins_b.pred_boxes_3d.R.numpy(),
                                               ins_b.projected_boxes.numpy())
        a.step(kf, ins_a, pose_np); b.step(kf, ins_b, pose_np)
        sa, sb = a.snapshot(), b.snapshot()
        for key in sa:
            assert sa[key].shape == sb[key].shape and np.array_equal(_bits(sa[key]), _bits(sb[key])), (k, key)
    assert a.count == len(kfs)


def test_degenerate_boxes():
    # thinnest boxes the pipeline allows (dims floor 0.01, box_fusion.py:719), nested boxes, touching boxes
    t = np.array([[0, 0, 0, 0.01, 0.01, 0.01], [0, 0, 0, 2.0, 2.0, 2.0], [0.004, 0, 0, 0.01, 0.01, 0.01],
                  [2.0, 0, 0, 2.0, 2.0, 2.0], [50, 50, 1, 1.0, 0.01, 1.0]], np.float32)
    R = np.tile(np.eye(3, dtype=np.float32), (5, 1, 1))
    c = ops.box_corners(t, R)
    iou, cnt = ops.iou3d_matrix(c, c, want_counts=True)
    ch = c.cpu().numpy()
    ia, ib = np.meshgrid(np.arange(5), np.arange(5), indexing="ij")
    gate, ref = port.obb_counts_pairs_c(ch, ia.ravel(), ib.ravel())
    assert np.array_equal(cnt.cpu().numpy().reshape(-1, 3), ref)
    assert np.array_equal(iou.cpu().numpy().ravel(), np.where(gate > 0, port.iou_from_counts(ref), 0.0))
    assert float(iou[0, 1]) > 0 and float(iou[1, 3]) > 0 and float(iou[0, 4]) == 0.0      # nested, touching faces, far apart


def test_fusion_list_capacity_is_reported():
    cfg = make_cfg("ca1m", pst_path=make_pst(32))
    bm = api.BoxManager(cfg)
    bm.fusion_list = [list(range(40))]
    bm.fusion_flag = [0]
    with pytest.raises(RuntimeError):
        bm.pack_lists(1)
