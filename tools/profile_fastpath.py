"""Host profile of the reference-shaped API on the engine-backed fast path (cProfile over the second half of the bench
sequence) and, with --engine, of FusionEngine.step.  Diagnostic."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                            # noqa: E402
from boxfusion_b200 import api, ops                                     # noqa: E402
from boxfusion_b200.driver import FusionSession                         # noqa: E402
from boxfusion_b200.engine import FusionEngine, pack_keyframe           # noqa: E402
from boxfusion_b200.synthetic import make_cfg                           # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg = make_cfg("ca1m", pst_path=bench.GOLDEN_PST, pst_size=1024)
    frames = bench.build_keyframes(1)
    for kf in frames:
        bench.pin_keyframe(kf)
        kf._resident = kf._pinned.to(dev)
    if "--engine" in sys.argv:
        eng = FusionEngine(cfg, device=dev)
        packed = [torch.from_numpy(pack_keyframe(k.tensor_cam, k.R_cam, k.scores, k.pred_boxes, k.pred_proj_xy, k.pose, k.K, k.image_size, i)).pin_memory()
                  for i, k in enumerate(frames)]
        for rep in range(2):
            eng.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for p, k in zip(packed, frames):
                eng.step(p, k.tensor_cam.shape[0])
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(f"engine: host issue {1e3 * (t1 - t0) / 300:.4f} ms/keyframe, until the GPU is done {1e3 * (t2 - t0) / 300:.4f} ms/keyframe")
        return
    sess = FusionSession(api, cfg, device=str(dev))
    pr = cProfile.Profile()
    for k, kf in enumerate(frames):
        if k == 100:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pr.enable()
        ins, pose_np = bench.make_instances(sess, kf, api, False)
        sess.step(kf, ins, pose_np)
    torch.cuda.synchronize()
    pr.disable()
    print(f"API: {1e3 * (time.perf_counter() - t0) / 200:.4f} ms/keyframe wall (under cProfile)")
    pstats.Stats(pr).sort_stats("tottime" if "--tottime" in sys.argv else "cumulative").print_stats(45)


if __name__ == "__main__":
    main()
