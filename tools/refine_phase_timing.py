"""Diagnostic: per-iteration phase cycle counts of bf_refine_kernel (leader CTA, thread 0), via BF_REFINE_TIMING=1.
Prints, per case, the mean SM cycles of {own evaluations, waiting for the rest of the cluster, leader phase, publish}."""
import json
import os
import sys

os.environ.setdefault("BF_REFINE_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np   # noqa: E402
import torch         # noqa: E402

from boxfusion_b200 import ops                                                     # noqa: E402
from boxfusion_b200.synthetic import make_cfg, make_pst, refine_problem            # noqa: E402


def case(name, B, V, P):
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
    rcfg = ops.make_refine_cfg(cfg, K16.reshape(-1), H, W)
    for _ in range(3):
        out = ops.refine(pst, t, R, s, uv, po, off, idx, rcfg, want_trace=True, max_views=V)
    torch.cuda.synchronize()
    its = out[2].cpu().numpy()
    tr = out[3].cpu().numpy().reshape(B, -1, 8)
    rows = np.concatenate([tr[b, :its[b], 1:8] for b in range(B)], 0)[:, [1, 2, 3, 4, 5, 6, 0]]
    m = rows.mean(0)
    print(json.dumps({"case": name, "launch": ops.last_refine_launch() if hasattr(ops, "last_refine_launch") else None,
                      "iters_mean": float(its.mean()),
                      "cycles_mean": {"own_evals": float(m[0]), "wait_cluster": float(m[1]), "leader": float(m[2]), "publish": float(m[3]), "leader_select": float(m[4]), "warp_eval_max": float(m[5]), "warp_eval_mean": float(m[6])},
                      "cycles_p90": [float(x) for x in np.percentile(rows, 90, axis=0)]}))


if __name__ == "__main__":
    case("C2-like 7x6x1024", 7, 6, 1024)
    case("C1 35x8x512", 35, 8, 512)
    case("3x4x1024", 3, 4, 1024)
    case("C4-ish 128x32x4096", 128, 32, 4096)
