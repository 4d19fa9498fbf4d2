#!/usr/bin/env python
"""Kernel-level stress measurements for the BASELINE.json configs that are not the bench line:
  C1  one fusion step: 50 detections vs 200-box map (N=250) NMS + refine of 35 boxes x 8 views, P=512
  C3  256 x 4096 IoU matrix (SAMPLED_REF and ANALYTIC) + 3-D NMS over N=4352
  C4  4096 particles x 32 views x 128 boxes, 20 forced iterations and early-stop run
CUDA events on the launching stream, 3 warm-ups, L2 flushed between repetitions.  Prints one JSON line per case."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from boxfusion_b200 import ops                                                  # noqa: E402
from boxfusion_b200.synthetic import make_cfg, make_pst, map_and_detections, refine_problem  # noqa: E402

FLOP_PER_EVAL = 1600.0


def timeit(fn, reps=5, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(min(ms))


def refine_case(name, B, V, P, early_stop, peak):
    prob = refine_problem(B, V, seed=11)
    W, H = prob["size"]
    pst = torch.from_numpy(make_pst(P, seed=1)).cuda()
    cfg = make_cfg("ca1m", pst_path=None, pst_size=P)
    K16 = np.eye(4, dtype=np.float32); K16[:3, :3] = prob["K"]
    dev = "cuda"
    t = torch.from_numpy(prob["tensor"].reshape(-1, 6)).to(dev); R = torch.from_numpy(prob["R"].reshape(-1, 9)).to(dev)
    s = torch.from_numpy(prob["scores"].reshape(-1)).to(dev); po = torch.from_numpy(prob["poses"].reshape(-1, 16)).to(dev)
    corners = ops.box_corners(t, R)
    uv = ops.project_boxes(corners, torch.linalg.inv(po.reshape(-1, 4, 4)), prob["K"], W, H).reshape(-1, 16)
    off = torch.arange(B + 1, dtype=torch.int32, device=dev) * V
    idx = torch.arange(B * V, dtype=torch.int32, device=dev)
    rcfg = ops.make_refine_cfg(cfg, K16.reshape(-1), H, W, early_stop=early_stop)
    res = {}

    def run():
        res["out"] = ops.refine(pst, t, R, s, uv, po, off, idx, rcfg, max_views=V)   # like BoxFusion.boxfusion / FusionEngine
    med, best = timeit(run)
    its = res["out"][2].cpu().numpy()
    evals = float(its.sum()) * P * V
    print(json.dumps({"case": name, "B": B, "V": V, "P": P, "early_stop": early_stop, "launch": ops.last_refine_launch(), "ms": round(med, 4), "ms_best": round(best, 4),
                      "iters_mean": round(float(its.mean()), 2), "evals": evals, "evals_per_s": round(evals / (med * 1e-3), 1),
                      "ms_per_iteration": round(med / float(its.max()), 4),
                      "fp32_tflops_algorithmic": round(evals * FLOP_PER_EVAL / (med * 1e-3) / 1e12, 3),
                      "frac_of_measured_fp32_peak": round(evals * FLOP_PER_EVAL / (med * 1e-3) / 1e12 / peak, 4)}))


def iou_case(peak):
    (mt, mR, ms_), (dt, dR, ds) = map_and_detections(4096, 256, seed=3, tilt_noise=0.0)
    ca = ops.box_corners(dt, dR); cb = ops.box_corners(mt, mR)
    for mode, nm in ((ops.IOU_SAMPLED_REF, "SAMPLED_REF"), (ops.IOU_ANALYTIC, "ANALYTIC")):
        res = {}

        def run():
            res["o"] = ops.iou3d_matrix(ca, cb, mode=mode, want_stats=True)
        med, best = timeit(run)
        st = res["o"][1].cpu().numpy()
        print(json.dumps({"case": "C3 iou matrix 256x4096", "mode": nm, "ms": round(med, 4), "pairs": int(st[0]),
                          "pairs_per_s": round(st[0] / (med * 1e-3), 1), "aabb_pass": int(st[1]), "gate_pass": int(st[2]),
                          "analytic": int(st[3]), "gate_pass_fraction": round(float(st[2]) / float(st[0]), 6)}))
    # NMS over N = 4352 (map + detections), fresh singleton fusion lists
    t = np.concatenate([mt, dt]); R = np.concatenate([mR, dR]); sc = np.concatenate([ms_, ds])
    n = t.shape[0]
    corners, centers = ops.box_corners(t, R, want_centers=True)
    order = torch.argsort(torch.from_numpy(sc).cuda(), descending=True, stable=True).to(torch.int32)
    iid = torch.arange(n, dtype=torch.int32, device="cuda")
    poses = torch.eye(4, device="cuda").reshape(1, 16).repeat(n, 1).contiguous()
    for mode, nm in ((ops.IOU_SAMPLED_REF, "SAMPLED_REF"), (ops.IOU_ANALYTIC, "ANALYTIC")):
        res = {}

        def run():
            fl = torch.zeros((n, ops.FUSION_CAP), dtype=torch.int32, device="cuda"); fl[:, 0] = iid
            ln = torch.ones(n, dtype=torch.int32, device="cuda"); fg = torch.zeros(n, dtype=torch.int32, device="cuda")
            res["o"] = ops.nms3d(corners, centers, order, iid, poses, fl, ln, fg, 0.1, 0.8, 30.0, 0.5, mode)
        med, best = timeit(run)
        keep = int(res["o"][0].sum().item())
        print(json.dumps({"case": "C3 nms N=4352", "mode": nm, "ms": round(med, 4), "pairs": n * (n - 1) // 2,
                          "pairs_per_s": round(n * (n - 1) / 2 / (med * 1e-3), 1), "kept": keep}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="c1,c3,c4")
    args = ap.parse_args()
    peak = ops.probe_fp32()
    print(json.dumps({"fp32_fma_peak_tflops_measured": round(peak, 2)}))
    cases = args.cases.split(",")
    if "c2" in cases:
        refine_case("C2-like refine launch (7 boxes x 6 views x 1024 particles)", 7, 6, 1024, True, peak)
    if "c1" in cases:
        refine_case("C1 refine (35 boxes x 8 views x 512 particles)", 35, 8, 512, True, peak)
    if "c4f" in cases:    # forced-iteration C4 only (tuning sweeps)
        refine_case("C4 refine forced 20 iterations", 128, 32, 4096, False, peak)
    if "c4" in cases:
        refine_case("C4 refine forced 20 iterations", 128, 32, 4096, False, peak)
        refine_case("C4 refine early stop", 128, 32, 4096, True, peak)
    if "prof" in cases:     # short launch for `ncu --set full` (one warm-up + one timed repetition are enough there)
        refine_case("profile refine (148 boxes x 16 views x 2048 particles, 20 iterations)", 148, 16, 2048, False, peak)
    if "c3" in cases:
        iou_case(peak)


if __name__ == "__main__":
    main()
