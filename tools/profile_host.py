"""cProfile of the host side of the per-keyframe step (run on the GPU box)."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                # noqa: E402
from boxfusion_b200 import api, ops                         # noqa: E402
from boxfusion_b200.driver import FusionSession             # noqa: E402
from boxfusion_b200.synthetic import make_cfg               # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
frames = bench.build_keyframes(1, n)
dev = torch.device("cuda", 0)
for kf in frames:
    bench.pin_keyframe(kf)
cfg = make_cfg("ca1m", pst_path=bench.GOLDEN_PST, pst_size=1024)


def run():
    sess = FusionSession(api, cfg, device="cuda:0")
    for kf in frames:
        ins, pose_np = bench.make_instances(sess, kf, api, False)
        sess.step(kf, ins, pose_np)
    torch.cuda.synchronize()


run()
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(25)
