#!/usr/bin/env python
"""Randomised stress parity (GPU): many seeds / shapes / tilt levels, every keyframe:
   engine == reference-shaped CUDA API == CPU port (C backend).  Reports how often the rarely-taken branches fired."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from boxfusion_b200 import api                                          # noqa: E402
from boxfusion_b200.driver import FusionSession                         # noqa: E402
from boxfusion_b200.engine import FusionEngine, pack_keyframe           # noqa: E402
from boxfusion_b200.synthetic import SyntheticScene, make_cfg, make_pst  # noqa: E402
from oracle import port                                                 # noqa: E402

KEYS = ("tensor", "R", "scores", "valid_num", "init_id", "fusion_flat", "fusion_off", "fusion_flag", "already_flat", "already_off")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8) if a.dtype.kind == "f" else a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=12)
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--with-port", type=int, default=4, help="first K seeds are also checked against the CPU port")
    ap.add_argument("--seed-base", type=int, default=300, help="scene seed of the first run (another base = another set of sequences)")
    args = ap.parse_args()
    port.IOU_BACKEND = "c_batch"
    stats = {"keyframes": 0, "keyframes_on_engine_backed_api": 0, "swaps_or_merges": 0, "fused": 0, "mismatch": 0}
    for seed in range(args.seeds):
        shape = "scannet" if seed % 2 else "ca1m"
        tilt = (0.0, 0.01, 0.03)[seed % 3]
        n_obj = (40, 90, 160)[seed % 3]
        scene = SyntheticScene(n_objects=n_obj, seed=args.seed_base + seed, max_det=(20, 35, 50)[seed % 3], shape=shape, tilt_noise=tilt,
                               new_frac=(0.1, 0.25)[seed % 2])
        P = (128, 256, 500)[seed % 3]
        cfg = make_cfg(shape, pst_path=make_pst(512, seed=seed), pst_size=P)
        eng = FusionEngine(cfg, map_capacity=2048, store_capacity=8192, fused_capacity=4096)
        sess = FusionSession(api, cfg, device="cuda")
        ref = FusionSession(port, cfg) if seed < args.with_port else None
        prev_valid = 0.0
        for k in range(args.frames):
            kf = scene.keyframe(k)
            eng.step(pack_keyframe(kf.tensor_cam, kf.R_cam, kf.scores, kf.pred_boxes, kf.pred_proj_xy, kf.pose),
                     kf.tensor_cam.shape[0], kf.K, kf.image_size)
            stats["keyframes_on_engine_backed_api"] += int(sess.box_manager._session is not None)
            sess.step(kf)
            a, b = eng.snapshot(), sess.snapshot()
            c = None
            if ref is not None:
                ref.step(kf)
                c = ref.snapshot()
            for key in KEYS:
                ok = a[key].shape == b[key].shape and np.array_equal(bits(a[key]), bits(b[key]))
                if c is not None:
                    ok = ok and c[key].shape == b[key].shape and np.array_equal(bits(c[key]), bits(b[key]))
                if not ok:
                    stats["mismatch"] += 1
                    print("MISMATCH seed", seed, "frame", k, key)
            stats["keyframes"] += 1
        stats["fused"] += len(sess.box_manager.already_fusion)
        stats["swaps_or_merges"] += sum(1 for l in sess.box_manager.fusion_list if len(l) > 5)
        print(f"seed {seed}: {shape} tilt {tilt} map {len(sess.all_pred_box)} fused {len(sess.box_manager.already_fusion)} "
              f"max list {max(len(l) for l in sess.box_manager.fusion_list)}", flush=True)
    print(stats)
    sys.exit(1 if stats["mismatch"] else 0)


if __name__ == "__main__":
    main()
