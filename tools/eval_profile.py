"""Diagnostic: cycles per section of one (particle, view) evaluation (bf_debug.cu, BF_EVAL_PROFILE ticks in bf_refine_eval.cuh).
Prints, per case and instantiation, the mean cycles a warp spends in each section of an evaluation."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np   # noqa: E402
import torch         # noqa: E402

from boxfusion_b200 import ops                                                     # noqa: E402
from boxfusion_b200.ops import ptr                                                 # noqa: E402
from boxfusion_b200.synthetic import make_cfg, make_pst, refine_problem            # noqa: E402

SECTIONS = ["corners", "project", "sort8", "chain8", "classify", "A_in_B", "gather+B_in_A", "pairs", "sortN", "chainN", "shoelace", "store"]


def run(name, V, P, C, T, roll, search, dist=2.5, reps=8, seed=11):
    prob = refine_problem(1, V, seed=seed)
    W, H = prob["size"]
    dev = "cuda"
    pst = torch.from_numpy(make_pst(P, seed=1)).to(dev)
    t = torch.from_numpy(prob["tensor"].reshape(-1, 6)).to(dev); R = torch.from_numpy(prob["R"].reshape(-1, 9)).to(dev)
    po_np = prob["poses"].reshape(-1, 4, 4).copy()
    if dist != 2.5:                                              # move the cameras towards the box: the box is cut by the image border
        c = prob["tensor"].reshape(-1, 6)[:, :3].mean(0)
        po_np[:, :3, 3] = c + (po_np[:, :3, 3] - c) * (dist / 2.5)
    po = torch.from_numpy(po_np.reshape(-1, 16)).to(dev)
    corners = ops.box_corners(t, R)
    uv = ops.project_boxes(corners, torch.linalg.inv(po.reshape(-1, 4, 4)), prob["K"], W, H).reshape(-1, 16)
    clamped = float(((uv <= 0) | (uv.reshape(-1, 8, 2) >= torch.tensor([W, H], device=dev)).reshape(-1, 16)).float().mean())
    state = torch.cat([t[0], torch.tensor(search, dtype=torch.float32, device=dev), R[0]]).contiguous()
    K = prob["K"]
    intr = torch.tensor([K[0, 0], K[1, 1], K[0, 2], K[1, 2], W, H], dtype=torch.float32, device=dev)
    PB = ((P + C - 1) // C + 31) & ~31
    out = torch.zeros(C * PB * V, dtype=torch.float32, device=dev)
    cyc = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
    h = ops.handle(torch.device(dev))
    fn = h.lib.bf_debug_eval_profile
    grid = min(C, 64 // (T // 32))
    for _ in range(2):
        rc = fn(ptr(pst), P, PB, ptr(state), ptr(po), ptr(uv), V, ptr(intr), grid, T, int(roll), reps, ptr(out), ptr(cyc), h.stream())
        assert rc == 0, rc
    torch.cuda.synchronize()
    c = cyc.cpu().numpy().reshape(64, 16)[: grid * (T // 32), :12].astype(np.float64)
    items = PB * V
    busy = c.sum(1) > 0
    passes = np.maximum(1, -(-items // T))
    per_eval = c[busy] / (reps * passes)                         # a busy warp makes `passes` evaluations per repetition (last may be partial)
    m = per_eval.mean(0)
    print(json.dumps({"case": name, "V": V, "P": P, "C": C, "T": T, "roll": bool(roll), "clamped_frac": round(clamped, 3),
                      "warps_busy": int(busy.sum()), "evals_per_warp_pass": int(passes), "mean_iou_term": float(out[: items].mean()),
                      "cycles_per_eval": {k: round(float(x)) for k, x in zip(SECTIONS, m)}, "total": round(float(m.sum())),
                      "slowest_warp": round(float(per_eval.sum(1).max())),
                      "slowest_sections": {k: round(float(x)) for k, x in zip(SECTIONS, per_eval[int(per_eval.sum(1).argmax())])},
                      "warp_totals": [round(float(x)) for x in per_eval.sum(1)]}))


if __name__ == "__main__":
    early, late = [0.1] * 3 + [0.5] * 3, [0.01] * 3 + [0.03] * 3
    for roll in (0, 1):
        run("early search, box in view", 6, 1024, 16, 384, roll, early)
        run("late search, box in view", 6, 1024, 16, 384, roll, late)
        run("late search, box cut by the border", 6, 1024, 16, 384, roll, late, dist=1.2)
        run("one warp alone", 1, 32, 1, 32, roll, late)
        run("bench shape 3 views T=512", 3, 1024, 16, 512, roll, late)
