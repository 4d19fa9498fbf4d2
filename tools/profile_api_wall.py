"""Wall-clock split of one keyframe on the reference-shaped API (engine-backed fast path), without a profiler: perf_counter around
the calls demo.py makes.  Diagnostic."""
import os
import sys
import time
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                            # noqa: E402
from boxfusion_b200 import api, fastpath                                # noqa: E402
from boxfusion_b200.driver import FusionSession                         # noqa: E402
from boxfusion_b200.synthetic import make_cfg                           # noqa: E402

T = defaultdict(float)
N = defaultdict(int)


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            T[name] += time.perf_counter() - t0
            N[name] += 1
    return w


def main():
    dev = torch.device("cuda", 0)
    cfg = make_cfg("ca1m", pst_path=bench.GOLDEN_PST, pst_size=1024)
    frames = bench.build_keyframes(1)
    for kf in frames:
        bench.pin_keyframe(kf)
        kf._resident = kf._pinned.to(dev)
    sess = FusionSession(api, cfg, device=str(dev))
    I = api.Instances3D
    I.cat = staticmethod(timed("cat", I.cat))
    I.spatial_association = staticmethod(timed("spatial_association", I.spatial_association))
    I.correspondence_association = staticmethod(timed("correspondence_association", I.correspondence_association))
    I.project_3d_boxes = timed("  project_3d_boxes", I.project_3d_boxes)
    api.GeneralInstance3DBoxes.transform2world = timed("  transform2world", api.GeneralInstance3DBoxes.transform2world)
    bm = sess.box_manager
    bm.update = timed("update", bm.update)
    bm.check_valid_num = timed("check_valid_num", bm.check_valid_num)
    bm.init_new_predictions = timed("init_new_predictions", bm.init_new_predictions)
    sess.box_fuser.boxfusion = timed("boxfusion", sess.box_fuser.boxfusion)
    fastpath.Session._wait_flags = timed("  _wait_flags", fastpath.Session._wait_flags)
    fastpath.Session._run_ahead = timed("  _run_ahead", fastpath.Session._run_ahead)
    mk = timed("make_instances", bench.make_instances)
    step = timed("driver.step", sess.step)
    for k, kf in enumerate(frames):
        if k == 100:
            torch.cuda.synchronize()
            T.clear(); N.clear()
            t0 = time.perf_counter()
        ins, pose_np = mk(sess, kf, api, False)
        step(kf, ins, pose_np)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print(f"wall {1e3 * total / 200:.4f} ms/keyframe")
    for k in sorted(T, key=lambda k: -T[k]):
        print(f"{k:32s} {1e3 * T[k] / 200:8.4f} ms/keyframe  ({N[k]} calls)")


if __name__ == "__main__":
    main()
