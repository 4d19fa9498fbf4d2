"""cProfile of the host side of FusionEngine.step over one C2 sequence (run on the GPU box)."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                # noqa: E402
from boxfusion_b200.engine import FusionEngine, pack_keyframe   # noqa: E402
from boxfusion_b200.synthetic import make_cfg               # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
frames = bench.build_keyframes(1, n)
cfg = make_cfg("ca1m", pst_path=bench.GOLDEN_PST, pst_size=1024)
data = [(torch.from_numpy(pack_keyframe(k.tensor_cam, k.R_cam, k.scores, k.pred_boxes, k.pred_proj_xy, k.pose)).pin_memory(),
         k.tensor_cam.shape[0], k.K, k.image_size) for k in frames]


def run():
    eng = FusionEngine(cfg, device=torch.device("cuda", 0))
    for packed, nn, K, size in data:
        eng.step(packed, nn, K, size)
    eng.check_status()
    torch.cuda.synchronize()


run()
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(30)
st.sort_stats("tottime").print_stats(18)
