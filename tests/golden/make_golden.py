"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run only in the build container (needs /root/reference):  python tests/golden/make_golden.py

The reference ships no tests or golden vectors for the fusion path, so
parity is pinned by executing the reference itself (oracle/ref_harness.py: import shims + its CUDA
kernel string compiled verbatim for the host) on deterministic synthetic inputs
(boxfusion_b200/synthetic.py) and freezing inputs and outputs here.  Environment of record is
written to golden_meta.json (numpy / scipy / torch versions matter: SURVEY F8).
"""
import io
import contextlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                      # noqa: E402
from boxfusion_b200.synthetic import (SyntheticScene, make_cfg, map_and_detections,     # noqa: E402
                                      refine_problem)
from boxfusion_b200.driver import FusionSession           # noqa: E402

REF_PST = os.path.join(rh.REFERENCE_ROOT, "data", "pst_1024_0.tiff")

SEQUENCES = {
    "seq_ca1m": dict(n_objects=40, seed=1, max_det=20, shape="ca1m", tilt_noise=0.0, frames=12),
    "seq_scannet_tilt": dict(n_objects=40, seed=2, max_det=20, shape="scannet", tilt_noise=0.02, frames=10),
}
# a keyframe every 3rd frame (cfg data.gap = 3 like config/ca1m.yaml's 20 / scannet.yaml's 25 in spirit) with
# BoxManager.check_valid_num switched on: frame ids and the `count - gap` threshold are in FRAMES (demo.py:217, 297-298)
GAP_SEQUENCE = ("seq_gap3_valid", dict(n_objects=40, seed=21, max_det=12, shape="scannet", new_frac=0.35, frames=10), 3)


def gen_pst(ref):
    import cv2
    pst = np.ascontiguousarray(cv2.imread(REF_PST, -1))
    assert pst.shape == (1024, 6) and pst.dtype == np.float32
    np.save(os.path.join(HERE, "pst_1024_0.npy"), pst)
    return pst


def gen_sequence(ref, name, spec, gap=1):
    spec = dict(spec)
    frames = spec.pop("frames")
    scene = SyntheticScene(**spec)
    cfg = make_cfg(spec["shape"], pst_path=REF_PST, pst_size=1024)
    if gap > 1:
        cfg["data"]["gap"] = gap
        cfg["box_fusion"]["check_valid"] = True
    sess = FusionSession(ref, cfg, frame_stride=gap)
    out = {"n_frames": np.array(frames)}
    for k in range(frames):
        kf = scene.keyframe(k)
        with contextlib.redirect_stdout(io.StringIO()):
            ins, pose_np = sess.make_pred_instances(kf)       # demo.py:216-221 with the reference's own ops
        out[f"k{k}_tensor_w"] = ins.pred_boxes_3d.tensor.numpy().copy()
        out[f"k{k}_R_w"] = ins.pred_boxes_3d.R.numpy().copy()
        out[f"k{k}_projected"] = ins.projected_boxes.numpy().copy()
        sess.step(kf, ins, pose_np)
        for key, val in sess.snapshot().items():
            out[f"k{k}_snap_{key}"] = val
        out[f"k{k}_mask"] = np.asarray(sess.last_mask if sess.last_mask is not None else [], dtype=np.int64)
        out[f"k{k}_success"] = np.asarray(sess.last_success if sess.last_success is not None else [], dtype=np.int64)
        sess.last_mask = sess.last_success = None
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "map", len(sess.all_pred_box), "fused", len(sess.box_manager.already_fusion))


def gen_iou_pairs(ref):
    (mt, mR, ms), (dt, dR, ds) = map_and_detections(60, 24, seed=11, tilt_noise=0.01)
    t = np.concatenate([mt, dt]); R = np.concatenate([mR, dR])
    corners = ref.GeneralInstance3DBoxes(torch.from_numpy(t), torch.from_numpy(R)).corners.numpy()
    n = t.shape[0]
    ia, ib = np.triu_indices(n, 1)
    iou = np.array([ref.Instances3D.obb_iou(corners[a], corners[b]) for a, b in zip(ia, ib)], dtype=np.float64)
    # 240 deliberately overlapping pairs (a box and a noisy re-observation) so the sampled branch
    # (instances.py:585-608) is well covered
    from boxfusion_b200.synthetic import random_boxes, box_rotation
    rs = np.random.RandomState(5)
    bt, bR = random_boxes(240, 21, tilt_noise=0.01)
    yaw = np.arctan2(bR[:, 1, 0], bR[:, 0, 0]) + rs.normal(0, 0.15, 240)
    ot = np.concatenate([bt[:, :3] + rs.normal(0, 0.12, (240, 3)), bt[:, 3:] * np.exp(rs.normal(0, 0.2, (240, 3)))], 1)
    oR = box_rotation(yaw, rs.normal(0, 0.01, 240), rs.normal(0, 0.01, 240)).astype(np.float32)
    ca = ref.GeneralInstance3DBoxes(torch.from_numpy(bt), torch.from_numpy(bR)).corners.numpy()
    cb = ref.GeneralInstance3DBoxes(torch.from_numpy(ot.astype(np.float32)), torch.from_numpy(oR)).corners.numpy()
    iou2 = np.array([ref.Instances3D.obb_iou(ca[i], cb[i]) for i in range(240)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "iou_pairs.npz"), tensor=t, R=R, corners=corners,
                        ia=ia.astype(np.int32), ib=ib.astype(np.int32), iou=iou,
                        ov_a=ca, ov_b=cb, ov_iou=iou2)
    print("overlap pairs", len(iou2), "nonzero", int((iou2 > 0).sum()))
    print("iou_pairs", len(iou), "nonzero", int((iou > 0).sum()), "over thr", int((iou > 0.1).sum()))


def gen_refine(ref, pst):
    out = {}
    cfg = make_cfg("ca1m", pst_path=REF_PST, pst_size=1024)
    for V in (3, 5, 8):
        prob = refine_problem(6, V, seed=100 + V)
        W, H = prob["size"]
        bf = ref.BoxFusion(cfg)
        bf.update_intrinsics((W, H), prob["K"])
        B = prob["tensor"].shape[0]
        per = ref.Instances3D((H, W))
        per.pred_boxes_3d = ref.GeneralInstance3DBoxes(torch.from_numpy(prob["tensor"].reshape(-1, 6)),
                                                       torch.from_numpy(prob["R"].reshape(-1, 3, 3)))
        per.cam_pose = torch.from_numpy(prob["poses"].reshape(-1, 4, 4))
        per.scores = torch.from_numpy(prob["scores"].reshape(-1))
        per.pred_boxes = torch.zeros(B * V, 4)
        per.project_3d_boxes(prob["K"], H=H, W=W)
        proj = per.projected_boxes.numpy().reshape(B, V, 8, 2)
        search = np.array([0.1, 0.1, 0.1, 0.5, 0.5, 0.5], np.float32)
        fit = np.stack([bf.evaluate_iou(prob["tensor"][b, 0].astype(np.float64), proj[b], prob["R"][b, 0],
                                        prob["scores"][b], prob["poses"][b], search, V) for b in range(B)])
        allp = ref.Instances3D((H, W))
        allp.pred_boxes_3d = ref.GeneralInstance3DBoxes(torch.from_numpy(prob["tensor"][:, 0].copy()),
                                                        torch.from_numpy(prob["R"][:, 0].copy()))
        bm = ref.BoxManager(cfg)
        bm.fusion_list = [list(range(b * V, (b + 1) * V)) for b in range(B)]
        bm.fusion_flag = [0] * B
        with contextlib.redirect_stdout(io.StringIO()):
            bf.boxfusion(allp, per, bm)
        for k in ("tensor", "R", "scores", "poses", "K"):
            out[f"v{V}_{k}"] = prob[k]
        out[f"v{V}_size"] = np.array(prob["size"])
        out[f"v{V}_projected"] = proj
        out[f"v{V}_fitness0"] = fit
        out[f"v{V}_fused"] = allp.pred_boxes_3d.tensor.numpy().copy()
        out[f"v{V}_flag"] = np.array(bm.fusion_flag)
        print("refine V", V, "fused", sum(bm.fusion_flag))
    np.savez_compressed(os.path.join(HERE, "refine_cases.npz"), **out)


def gen_hull_helpers(ref):
    """The small public helpers around obb_iou (instances.py:493-571, 616-641) and init_opt_params_v2
    (box_fusion.py:602-619), run on the pairs of iou_pairs.npz: `python tests/golden/make_golden.py helpers` adds this
    fixture without touching the others."""
    g = np.load(os.path.join(HERE, "iou_pairs.npz"))
    corners = g["corners"]
    rs = np.random.RandomState(5)
    sel = rs.choice(len(g["ia"]), 600, replace=False)
    pa = np.concatenate([corners[g["ia"][sel]], g["ov_a"]]).astype(np.float32)
    pb = np.concatenate([corners[g["ib"][sel]], g["ov_b"]]).astype(np.float32)
    gate = np.array([ref.Instances3D.check_intersection(a, b) for a, b in zip(pa, pb)])
    pts, inside = [], []
    for a, b in zip(g["ov_a"][:80], g["ov_b"][:80]):
        lo, hi = np.minimum(a.min(0), b.min(0)), np.maximum(a.max(0), b.max(0))
        p = np.concatenate([rs.uniform(lo, hi, (40, 3)), ref.Instances3D.augment_vertices(a.astype(np.float32)).astype(np.float64),
                            a.astype(np.float64)])          # random points, the partner's 20 gate points, own corners (on the hull)
        pts.append(p); inside.append(ref.Instances3D.batch_in_convex_hull_3d(p, b))
    aug = np.stack([ref.Instances3D.augment_vertices(c) for c in corners[:6]])
    A2 = rs.uniform(0, 300, (20, 8, 2)).astype(np.float32)
    B2 = np.sort(rs.uniform(0, 300, (20, 7, 2, 2)), axis=2).reshape(20, 7, 4)[..., [0, 2, 1, 3]]
    i2 = np.stack([np.stack(ref.Instances3D.IoU_2D(a, b)) for a, b in zip(A2, B2)])
    bf = ref.BoxFusion(make_cfg("ca1m", pst_path=REF_PST, pst_size=1024))
    vb = rs.uniform(0.2, 2.0, (9, 5, 6)); vs = rs.uniform(0, 1, (9, 5)); vR = rs.normal(0, 1, (9, 5, 3, 3))
    v2 = [bf.init_opt_params_v2(b, r, s_) for b, r, s_ in zip(vb, vR, vs)]
    np.savez_compressed(os.path.join(HERE, "hull_helpers.npz"), pa=pa, pb=pb, gate=gate, pts=np.stack(pts), pts_box=g["ov_b"][:80],
                        inside=np.stack(inside), aug_in=corners[:6], aug=aug, iou2d_A=A2, iou2d_B=B2, iou2d=i2,
                        v2_boxes=vb, v2_scores=vs, v2_R=vR, v2_mean=np.stack([m for m, _ in v2]), v2_rot=np.stack([r for _, r in v2]))
    print("hull_helpers: gate true for", int(gate.sum()), "of", len(gate), "pairs; inside", int(np.stack(inside).sum()), "of", np.stack(inside).size)


def gen_prefilters(ref):
    """The detection pre-filters (box_manager.py:217-245; demo.py:140-148) and the pose-disparity methods
    (box_manager.py:168-215) of the unmodified reference on seeded inputs: `python tests/golden/make_golden.py extras`."""
    rs = np.random.RandomState(0)
    n = 4000
    dims = np.exp(rs.normal(np.log(0.4), 1.0, (n, 3))).astype(np.float32)
    dims[:200, 0] = 3.0; dims[:200, 1] = 0.05                                  # floor-like slabs
    dims[200:300] = np.array([0.14, 0.13, 1.2], np.float32) * rs.uniform(0.9, 1.1, (100, 3)).astype(np.float32)
    dims[300:340] = np.array([1.5, 0.1, 0.1], np.float32)                      # exactly on the ratio thresholds (15, 7.5, 20, 10)
    dims[340:380] = np.array([2.0, 0.1, 0.1], np.float32)
    t = np.concatenate([rs.normal(0, 2, (n, 3)).astype(np.float32), dims], 1)
    uv = np.stack([rs.uniform(-20, 404, n), rs.uniform(-20, 532, n)], 1).astype(np.float32)
    uv[:50, 0] = np.array([38.0, 39.0, 345.0, 346.0, 64.0] * 10, np.float32)   # on / next to the integer bounds of both shapes
    uv[:50, 1] = np.array([51.0, 52.0, 460.0, 461.0, 48.0] * 10, np.float32)
    out = {"tensor": t, "uv": uv}
    tt, uu = torch.from_numpy(t), torch.from_numpy(uv)
    for shape in ("ca1m", "scannet"):
        cfg = make_cfg(shape)
        bm = ref.BoxManager(cfg)
        W, H = cfg["cam"]["W"], cfg["cam"]["H"]
        for ratio in (cfg["detection"]["uv_bound_value"], 1.0, 0.75):
            out[f"{shape}_uv_{ratio}"] = bm.check_uv_bounds(uu, W, H, ratio=ratio).numpy()
        for ratio in (cfg["detection"]["floor_ratio"], 20):
            out[f"{shape}_floor_{ratio}"] = bm.check_floor_mask(tt, ratio=ratio).numpy()
        for thres in (0.5, 2.5):
            out[f"{shape}_large_{thres}"] = bm.check_large_mask(tt, thres=thres).numpy()
    # pose disparity: random rigid poses incl. identical pairs, pure translations, 180-degree turns
    m = 600
    poses = np.tile(np.eye(4, dtype=np.float32), (m, 1, 1))
    for i in range(m):
        a = rs.normal(0, 1, (3, 3)); q, _ = np.linalg.qr(a)
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        poses[i, :3, :3] = q.astype(np.float32)
        poses[i, :3, 3] = rs.normal(0, 1.5, 3).astype(np.float32)
    poses[1] = poses[0]                                                         # identical
    poses[3, :3, :3] = poses[2, :3, :3]                                          # pure translation
    poses[5, :3, :3] = poses[4, :3, :3] @ np.diag([-1.0, -1.0, 1.0]).astype(np.float32)   # 180 degrees
    ia, ib = rs.randint(0, m, 1500), rs.randint(0, m, 1500)
    ia[:3], ib[:3] = [0, 2, 4], [1, 3, 5]
    centers = rs.normal(0, 1, (m, 3))
    bm = ref.BoxManager(make_cfg("ca1m"))
    P = torch.from_numpy(poses)
    res = np.zeros((1500, 4), np.float64)
    for k, (a, b) in enumerate(zip(ia, ib)):
        base, ang, score, cd = bm.compute_pose_center_disparity(P[a], P[b], centers[a], centers[b])
        b2, a2, s2 = bm.compute_pose_disparity(P[a], P[b])
        assert float(b2) == float(base) and (float(a2) == float(ang) or (np.isnan(float(a2)) and np.isnan(float(ang))))
        res[k] = (float(base), float(ang), float(score), float(cd))
    out.update(poses=poses, ia=ia, ib=ib, centers=centers, disparity=res)
    np.savez_compressed(os.path.join(HERE, "prefilters.npz"), **out)
    print("prefilters:", {k: int(v.sum()) for k, v in out.items() if v.dtype == np.bool_})


def gen_results(ref):
    """Result / wire formats: the reference's own tools/utils.py post_process / save_box (:302-332) on corner arrays, and the
    two save lists demo.py:369-387 builds from them (those five lines are transcribed here: demo.py itself cannot run)."""
    import pickle
    import tempfile
    tu = rh.load_reference_tools()
    rs = np.random.RandomState(3)
    (mt, mR, _), _ = map_and_detections(60, 8, seed=4, tilt_noise=0.01)
    mt[:20, 3:] *= rs.uniform(0.1, 0.6, (20, 3)).astype(np.float32)              # some boxes below the 0.3 m ScanNet size filter
    corners = ref.GeneralInstance3DBoxes(torch.from_numpy(mt), torch.from_numpy(mR)).corners.cpu().numpy()
    kept = tu.post_process(corners)                                               # tools/utils.py:302-317 (default threshold 0.3)
    kept05 = tu.post_process(corners, threshold=0.5)
    save_list = [[(int(0), (kept[n]), 1.0) for n in range(len(kept))]]            # demo.py:376-378 after post_process (scannet)
    feats = rs.normal(0, 1, (len(corners), 16)).astype(np.float32)
    cls = rs.randint(0, 30, len(corners))
    all_save_list = [[(cls[n], (corners[n]), feats[n]) for n in range(len(corners))]]     # demo.py:383-386
    d = tempfile.mkdtemp()
    with contextlib.redirect_stdout(io.StringIO()):
        tu.save_box(save_list, os.path.join(d, "a.pkl"))
        tu.save_box(all_save_list, os.path.join(d, "b.pkl"))
    back = tu.load_data(os.path.join(d, "a.pkl"))
    assert len(back[0]) == len(kept)
    np.savez_compressed(os.path.join(HERE, "results_formats.npz"), tensor=mt, R=mR, corners=corners, kept=kept, kept05=kept05,
                        features=feats, classes=cls,
                        global_pkl=np.frombuffer(open(os.path.join(d, "a.pkl"), "rb").read(), dtype=np.uint8),
                        framewise_pkl=np.frombuffer(open(os.path.join(d, "b.pkl"), "rb").read(), dtype=np.uint8),
                        pickle_protocol=np.array(pickle.HIGHEST_PROTOCOL))
    print("results_formats: post_process keeps", len(kept), "of", len(corners), "(0.5:", len(kept05), ")")


def main():
    ref = rh.load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "helpers":
        gen_hull_helpers(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "extras":       # round 2 fixtures, without touching the round-1 ones
        gen_prefilters(ref)
        gen_results(ref)
        gen_sequence(ref, GAP_SEQUENCE[0], GAP_SEQUENCE[1], gap=GAP_SEQUENCE[2])
        return
    pst = gen_pst(ref)
    for name, spec in SEQUENCES.items():
        gen_sequence(ref, name, spec)
    gen_iou_pairs(ref)
    gen_refine(ref, pst)
    gen_hull_helpers(ref)
    gen_prefilters(ref)
    gen_results(ref)
    gen_sequence(ref, GAP_SEQUENCE[0], GAP_SEQUENCE[1], gap=GAP_SEQUENCE[2])
    import scipy
    meta = {"numpy": np.__version__, "scipy": scipy.__version__, "torch": torch.__version__,
            "reference": "pliam1105/BoxFusion @ /root/reference", "sequences": SEQUENCES,
            "note": "reference kernel string compiled for the host with g++ -O2 -ffp-contract=off"}
    json.dump(meta, open(os.path.join(HERE, "golden_meta.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
