#!/usr/bin/env python
"""bench.py - BoxFusion multi-view box-fusion hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one keyframe of the synthetic CA-1M-shaped 300-keyframe sequence (BASELINE.json configs[1]): camera->world
lift + observation projection (A2/A15) + 3-D NMS association with fusion-list bookkeeping (A3-A8) + small-object
correspondence (A9-A12) + particle refinement of every fusable map box (A16-A22), made through the reference-shaped API
exactly as demo.py:200-327 calls it (boxfusion_b200/driver.py).  The K timed keyframes are every (300/K)-th keyframe of the
sequence - the map grows from 0 to its full size over the sequence, so a prefix would time the easy part only - and the
keyframes in between advance the state untimed.

Printed JSON line (rank 0):
  value        whole-job keyframes/s with every keyframe's detections already resident in HBM (reference-shaped API)
  e2e          the same with HOST (pinned) detections copied in and the call's results read back inside the timed region
  e2e_engine   the same keyframes through the engine's own entry (bf_engine_step: one C call per keyframe)
  roofline     dominant kernel (bf_refine_kernel): algorithmic FP32 flops / CUDA-event duration vs the FP32 FMA
               throughput measured on this device (bf_probe_fp32); the path is FP32-issue bound, not HBM or tensor
               (SURVEY.md section 8(d)); achieved HBM GB/s is reported beside it
  c1_step      BASELINE configs[0]: one captured fusion step (50 detections vs 200-box map, 35 boxes x 8 views x 512 particles)
  iou          BASELINE configs[2]: 256 x 4096 IoU matrix (SAMPLED_REF / ANALYTIC) + NMS over 4352 boxes, pairs/s
  c4           BASELINE configs[3]: 4096 particles x 32 views x 128 boxes
  c5           BASELINE configs[4]: 64 independent ScanNet-shaped sequences sharded over the ranks (strong scaling)
  cpu_baseline the CPU port of the reference algorithm (oracle/port.py, scipy/Qhull IoU + C kernel) on a bounded sample
Multi-GPU: independent sequences, one per rank (weak scaling), no data-path collective; a final NCCL all_gather
collects the per-rank maps.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from boxfusion_b200.synthetic import SyntheticScene, make_cfg          # noqa: E402
from boxfusion_b200.driver import FusionSession                         # noqa: E402

GOLDEN_PST = os.path.join(ROOT, "tests", "golden", "pst_1024_0.npy")
FRAMES_PER_SEQUENCE = 300
N_OBJECTS, MAX_DET = 200, 50
# algorithmic FP32 work of one (particle, view) evaluation of compute_iou_value, counted on the straight-line
# path of bf_eval_view for two hexagonal hulls with 6-8 intersection candidates (table in DESIGN.md section 4, K3)
FLOP_PER_EVAL = 1600.0
# bytes one evaluation must touch: nothing in HBM (PST row and view constants are on chip); 4 B of fitness per particle
BYTES_PER_EVAL = 0.5
# algorithmic work of the sampled IoU: 960 flop per pair (gate) + 750 000 per gate-passing pair
FLOP_PER_PAIR_GATE, FLOP_PER_PAIR_ESTIMATE, FLOP_PER_PAIR_ANALYTIC = 960.0, 750000.0, 700.0
METRIC = "fusion keyframes/s (= 1000 / fusion ms/frame), association + particle refine per keyframe"


def workload_config(K):
    """The `config` of the JSON line - identical for the product arm and the reference arm."""
    stride = FRAMES_PER_SEQUENCE / min(K, FRAMES_PER_SEQUENCE)
    return {"workload": "BASELINE configs[1]: synthetic CA-1M-shaped 300-keyframe sequence (384x512, 200 objects, <=50 detections/"
                        "keyframe, shipped 1024-particle template, 20 iters), one sequence per GPU; the timed keyframes are every "
                        f"{stride:g}-th keyframe of the sequence (the map grows over the sequence), the keyframes in between advance "
                        "the state untimed",
            "timed_keyframes_of_sequence": sample_indices(min(K, FRAMES_PER_SEQUENCE)),
            "iou_mode": "SAMPLED_REF (reference-exact)"}


def sample_indices(K):
    return [int(round((j + 1) * FRAMES_PER_SEQUENCE / K)) - 1 for j in range(K)]


def build_keyframes(seed: int, n: int = FRAMES_PER_SEQUENCE):
    sc = SyntheticScene(n_objects=N_OBJECTS, seed=seed, max_det=MAX_DET, shape="ca1m")
    return [sc.keyframe(k) for k in range(n)]


def plan(K, rank):
    """K timed steps -> [(keyframes of one sequence, set of timed indices)]."""
    out, left, s = [], K, 0
    while left > 0:
        k = min(left, FRAMES_PER_SEQUENCE)
        # every rank runs the same sequences (weak scaling: fixed work per GPU; with rank-dependent seeds the max over ranks
        # would time whichever rank drew the heaviest sequence)
        out.append((build_keyframes(17 * s + 1), set(sample_indices(k))))
        left -= k
        s += 1
    return out


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 100 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_FIELDS = (("tensor_cam", 6), ("R_cam", 9), ("scores", 1), ("pred_boxes", 4), ("pred_proj_xy", 2))


def pin_keyframe(kf):
    """Host staging of one keyframe's detections in ONE pinned buffer (field after field), so the e2e pass
    issues a single H2D copy per keyframe."""
    flat = np.concatenate([np.ascontiguousarray(getattr(kf, k), dtype=np.float32).reshape(-1) for k, _ in _FIELDS])
    kf._pinned = torch.from_numpy(flat).pin_memory()
    return flat.nbytes + 64   # + the 4x4 pose


def split_fields(buf, n):
    out, o = {}, 0
    for k, w in _FIELDS:
        t = buf[o:o + n * w]
        out[k] = t.view(n, 3, 3) if k == "R_cam" else (t if w == 1 else t.view(n, w))
        o += n * w
    return out


def make_instances(sess, kf, api, resident):
    """demo.py:216-221 for the CUDA product: the detector's output (pinned host or HBM-resident) -> Instances3D; the
    bookkeeping fields are host tensors exactly as demo.py:216-219 creates them."""
    from boxfusion_b200 import ops
    dev = sess.device
    n = kf.tensor_cam.shape[0]
    ins = api.Instances3D((kf.image_size[1], kf.image_size[0]))
    if resident:
        t = split_fields(kf._resident.clone(), n)
    else:
        t = split_fields(kf._pinned.to(dev, non_blocking=True), n)
        ops.Profile.h2d_bytes += kf._pinned.numel() * 4
    ins.scores, ins.pred_boxes, ins.pred_proj_xy = t["scores"], t["pred_boxes"], t["pred_proj_xy"]
    ins.pred_boxes_3d = api.GeneralInstance3DBoxes._wrap(t["tensor_cam"], t["R_cam"])
    pose_np = np.repeat(kf.pose[None], repeats=n, axis=0)
    ins.cam_pose = torch.from_numpy(pose_np)                                 # demo.py:216
    ins.frame_id = torch.tensor([sess.count]).repeat(n)                      # demo.py:217
    ins.init_id = sess.box_count + torch.arange(n)                           # demo.py:218
    ins.valid_num = torch.zeros(n)                                           # demo.py:219
    ins.pred_boxes_3d.transform2world(ins.cam_pose)                          # bf_transform2world (pose copy counted by ops)
    ins.project_3d_boxes(kf.K, H=kf.image_size[1], W=kf.image_size[0])       # bf_box_corners + bf_project_boxes
    return ins, pose_np


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run_ours(args, rank, world, local_rank):
    from boxfusion_b200 import _lib, api, ops
    from boxfusion_b200.engine import FusionEngine, pack_keyframe
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg = make_cfg("ca1m", pst_path=GOLDEN_PST, pst_size=1024)
    K, W = args.steps, max(args.warmup, 3)
    seqs = plan(K, rank)
    warm = build_keyframes(999 + rank, W + 12)
    for kf in [k for s, _ in seqs for k in s] + warm:
        pin_keyframe(kf)
        kf._resident = kf._pinned.to(dev)
        n = kf.tensor_cam.shape[0]
        kf._packed = torch.empty(_lib.KF_HEADER + _lib.KF_ROW * n, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    # ---- the reference-shaped API (driver.FusionSession = demo.py:200-327), timed keyframes bracketed by CUDA events ----
    def run_api(plan_, resident):
        ops.Profile.reset()
        evs, sess = [], None
        for frames, timed in plan_:
            sess = FusionSession(api, cfg, device=str(dev))
            for k, kf in enumerate(frames):
                if k in timed:
                    flush.zero_()                                       # L2 flush between steps, outside the step's events
                    a, b = _events()
                    a.record()
                    ins, pose_np = make_instances(sess, kf, api, resident)
                    sess.step(kf, ins, pose_np)
                    b.record()
                    evs.append((a, b))
                else:
                    ins, pose_np = make_instances(sess, kf, api, True)
                    sess.step(kf, ins, pose_np)
        torch.cuda.synchronize()
        return [x.elapsed_time(y) for x, y in evs], sess

    # ---- the engine's own entry: one C call per keyframe; host packing (incl. the two pose inverses) is timed ----
    def run_engine(plan_, phase_events=False):
        ops.Profile.reset()
        evs, fuse, eng = [], [], None
        for frames, timed in plan_:
            eng = FusionEngine(cfg, device=dev, map_capacity=4096, store_capacity=max(65536, 64 * len(frames)))
            for k, kf in enumerate(frames):
                n = kf.tensor_cam.shape[0]
                is_timed = k in timed
                if is_timed:
                    flush.zero_()
                    a, b = _events()
                    a.record()
                kf._packed.copy_(torch.from_numpy(pack_keyframe(kf.tensor_cam, kf.R_cam, kf.scores, kf.pred_boxes, kf.pred_proj_xy,
                                                               kf.pose, kf.K, kf.image_size, k)))
                if phase_events and is_timed and n:
                    # same keyframe issued phase by phase with events around the fuse phase (select + bf_refine + apply)
                    eng.step(kf._packed, n, phases=_lib.PH_INGEST | _lib.PH_NMS | _lib.PH_CORR | _lib.PH_COMPACT)
                    fa, fb = _events()
                    fa.record(); eng.step(kf._packed, n, phases=_lib.PH_FUSE); fb.record()
                    eng.step(kf._packed, n, phases=_lib.PH_FINISH)
                    s = eng.state()                                     # B / views of this keyframe (synchronises: measurement pass only)
                    fuse.append((fa, fb, int(s.B), int(s.SV), eng.refine_evals(int(s.B))))
                else:
                    eng.step(kf._packed, n)
                if is_timed:
                    b.record()
                    evs.append((a, b))
            eng.check_status()
        torch.cuda.synchronize()
        return [x.elapsed_time(y) for x, y in evs], eng, fuse

    warm_plan = [(warm, set(range(W)))]
    # warm-up (untimed): >= 3 steps of another sequence through every pass
    run_api(warm_plan, True)
    run_api(warm_plan, False)
    run_engine(warm_plan)
    run_engine(warm_plan, phase_events=True)
    fp32_peak = ops.probe_fp32()                                        # every rank, before the first barrier

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    if world > 1:                                                       # NCCL communicator set-up outside every timed region
        torch.distributed.all_reduce(torch.zeros(1, device=dev))
    # ---- pass 1: inputs resident in HBM -> `value` ------------------------------------------------
    sampler = ClockSampler(local_rank)
    barrier(); sampler.start(); t0 = time.perf_counter()
    step_ms, sess = run_api(seqs, True)
    barrier(); wall_resident = time.perf_counter() - t0
    launches, call_counts = ops.Profile.launches, dict(ops.Profile.calls)
    fast = sess.box_manager._session is not None
    # ---- pass 2: host inputs, H2D + D2H inside the timed region -> `e2e` -----------------------------------
    barrier(); t0 = time.perf_counter()
    step_ms_e2e, sess2 = run_api(seqs, False)
    barrier(); wall_e2e = time.perf_counter() - t0
    n_timed = len(step_ms_e2e)
    h2d_b, d2h_b = ops.Profile.h2d_bytes, ops.Profile.d2h_bytes
    # ---- context: the same API free-running (no L2 flush, no per-step events; one synchronise at the end): what a caller's loop
    #      over the sequence sees once the engine is warm - keyframes 20..299, host inputs, wall clock ----
    def run_api_free(frames, skip=20):
        s_ = FusionSession(api, cfg, device=str(dev))
        t_ = None
        for k, kf in enumerate(frames):
            if k == skip:
                torch.cuda.synchronize()
                t_ = time.perf_counter()
            ins, pose_np = make_instances(s_, kf, api, False)
            s_.step(kf, ins, pose_np)
        torch.cuda.synchronize()
        return (time.perf_counter() - t_) / max(len(frames) - skip, 1) if t_ is not None else None
    free_s = run_api_free(seqs[0][0]) if len(seqs[0][0]) > 40 else None
    # ---- context: the same pass with every call on its own (round 1's implementation; the engine-backed path switched off) ----
    from boxfusion_b200 import fastpath
    fastpath.ENABLED = False
    try:
        step_ms_cbc, sess3 = run_api(seqs, False)
    finally:
        fastpath.ENABLED = True
    cbc_h2d, cbc_d2h = ops.Profile.h2d_bytes, ops.Profile.d2h_bytes
    assert len(sess3.all_pred_box) == len(sess2.all_pred_box)
    # ---- pass 3: the engine's own entry, host inputs, same keyframes -----------------
    barrier(); t0 = time.perf_counter()
    step_ms_eng, eng, _ = run_engine(seqs)
    barrier(); wall_eng = time.perf_counter() - t0
    eng_h2d, eng_d2h, eng_launches = ops.Profile.h2d_bytes, ops.Profile.d2h_bytes, ops.Profile.launches
    n_all = sum(len(f) for f, _ in seqs)
    assert eng.N == len(sess.all_pred_box) == len(sess2.all_pred_box), "engine and API disagree on the final map size"
    # ---- pass 4 (measurement only): CUDA events around the fuse phase of every timed keyframe -> roofline ----------
    step_ms_roof, _, fuse = run_engine(seqs, phase_events=True)
    clocks = sampler.stop()

    t_res, t_e2e, t_eng = sum(step_ms) / 1e3, sum(step_ms_e2e) / 1e3, sum(step_ms_eng) / 1e3
    if world > 1:                                                       # max over ranks, on the device clock
        tt = torch.tensor([t_res, t_e2e, t_eng], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        t_res, t_e2e, t_eng = float(tt[0]), float(tt[1]), float(tt[2])
        # the only exchange of the job: gather every rank's final map (rows of 15 floats), SURVEY section 8(e)
        from boxfusion_b200.sharding import gather_maps, map_rows
        maps = gather_maps(map_rows(sess.all_pred_box))
        assert len(maps) == world
    if rank != 0:
        return None
    total_frames = K * world
    roof = None
    fz = [(a.elapsed_time(b), B, SV, ev) for a, b, B, SV, ev in fuse if B > 0]
    if fz:
        tot_ms, tot_ev = sum(x[0] for x in fz), float(sum(x[3] for x in fz))
        achieved = tot_ev * FLOP_PER_EVAL / (tot_ms * 1e-3) / 1e12
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_ach = tot_ev * BYTES_PER_EVAL / (tot_ms * 1e-3) / 1e9
        # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture of this command (current build)
        traffic, tsrc = None, None
        tpath = os.path.join(ROOT, "profiles", "r2_refine_bench_traffic.json")
        if os.path.isfile(tpath):
            tj = json.load(open(tpath))
            traffic, tsrc = round(tj["dram_bytes_per_launch_mean"]), tj["source"]
        roof = {"bound": "fp32", "kernel": "bf_refine_kernel", "achieved": round(achieved, 3), "peak": round(fp32_peak, 2),
                "unit": "TFLOP/s", "frac": round(achieved / fp32_peak, 4), "traffic": traffic, "traffic_unit": "bytes/launch",
                "traffic_source": tsrc, "algorithmic_bytes_per_launch": round(tot_ev * BYTES_PER_EVAL / len(fz)),
                "peak_source": "bf_probe_fp32 FMA micro-benchmark on this device (burst); MEASURED_PEAKS.json has no FP32 entry",
                "timed": "CUDA events around the fuse phase (bf_engine_select + bf_refine_kernel + bf_engine_apply; the two small "
                         "kernels are ~2 us each) of every timed keyframe that refined at least one box",
                "launches": len(fz), "avg_launch_ms": round(tot_ms / len(fz), 4),
                "boxes_per_launch": round(sum(x[1] for x in fz) / len(fz), 2), "views_per_launch": round(sum(x[2] for x in fz) / len(fz), 2),
                "evals_per_launch": round(tot_ev / len(fz), 1), "flop_per_eval": FLOP_PER_EVAL,
                "evals_per_s": round(tot_ev / (tot_ms * 1e-3), 1),
                "share_of_step": round(tot_ms / sum(step_ms_roof), 4),
                "hbm": {"achieved": round(hbm_ach, 3), "peak": hbm_peak, "unit": "GB/s", "frac": round(hbm_ach / hbm_peak, 6),
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    cfgd = workload_config(K)
    out = {
        "metric": METRIC,
        "value": round(total_frames / t_res, 3), "unit": "keyframes/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(1e3 * t_res / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfgd,
        "e2e": {"value": round(total_frames / t_e2e, 3), "unit": "keyframes/s", "ms_per_step": round(1e3 * t_e2e / K, 4),
                "h2d_bytes_per_step": int(h2d_b / max(n_all, 1)), "d2h_bytes_per_step": int(d2h_b / max(n_all, 1)),
                "api": "reference-shaped calls of demo.py:243-327 (Instances3D.cat / spatial_association / correspondence_association, "
                       "BoxManager.update, BoxFusion.boxfusion)" + (" on the engine-backed fast path" if fast else " (call by call)")},
        "e2e_engine": {"value": round(total_frames / t_eng, 3), "unit": "keyframes/s", "ms_per_step": round(1e3 * t_eng / K, 4),
                       "h2d_bytes_per_step": int(eng_h2d / max(n_all, 1)), "d2h_bytes_per_step": int(eng_d2h / max(n_all, 1)),
                       "gpu_launches": int(eng_launches), "launches_per_keyframe": eng.launch_counts[7],
                       "note": "same keyframes through bf_engine_step (one C call = one H2D copy + one CUDA-graph launch per keyframe; "
                               "host packing of the detections incl. both pose inverses is inside the timed region); final map "
                               "identical to the reference-shaped API's"},
        "e2e_free_running": (None if free_s is None else {
            "ms_per_step": round(1e3 * free_s, 4), "value": round(1.0 / free_s, 1), "unit": "keyframes/s", "rank": 0,
            "note": "context: the same reference-shaped API over keyframes 20..299 of the sequence in a plain loop (host inputs, no L2 flush, "
                    "no per-step events, one synchronise at the end): GPU work of one keyframe overlaps the host side of the next"}),
        "e2e_call_by_call": {"ms_per_step": round(sum(step_ms_cbc) / len(step_ms_cbc), 4), "h2d_bytes_per_step": int(cbc_h2d / max(n_all, 1)),
                             "d2h_bytes_per_step": int(cbc_d2h / max(n_all, 1)), "rank": 0,
                             "note": "context: the same keyframes with boxfusion_b200.fastpath.ENABLED = False - every reference-shaped call uploads, "
                                     "launches and downloads on its own (round 1's implementation of the API)"},
        "gpu_launches": int(launches), "calls": call_counts, "l2": "flushed between steps (256 MiB memset outside the step events)",
        "final_map_boxes": len(sess.all_pred_box), "keyframes_run": n_all,
        "refine_division_redos": ops.cold_redos(dev, reset=False),      # evaluations redone with plain divisions (bf_fdiv window): expected 0
        "wall_s": {"resident": round(wall_resident, 3), "e2e": round(wall_e2e, 3), "engine": round(wall_eng, 3)},
        "p50_ms": round(float(np.percentile(step_ms, 50)), 4), "p99_ms": round(float(np.percentile(step_ms, 99)), 4),
        "roofline": roof, "clocks": clocks,
    }
    return out, seqs, step_ms_e2e


# =====================================================================================================================
# the other BASELINE configs (each returns a dict for the JSON line)
# =====================================================================================================================
def _time_local(fn, dev, reps, warmup=3, restore=None):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        if restore:
            restore()
        fn()
    ms = []
    for _ in range(reps):
        if restore:
            restore()
        flush.zero_()
        torch.cuda.synchronize()
        a, b = _events()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(min(ms))


def block_c1(dev, fp32_peak):
    """BASELINE configs[0]: ONE fusion step on a CA-1M-shaped frame - 50 detections against a 200-box map (N = 250 through
    NMS + correspondence), 35 map boxes with 8 views each to refine, 512 particles - as one captured CUDA-graph launch
    (bf_engine_step); the engine state is restored before every repetition.  Median of 100."""
    from boxfusion_b200 import api, fastpath, ops
    from boxfusion_b200.engine import pack_keyframe
    from boxfusion_b200.synthetic import make_pst, map_and_detections, refine_problem
    B, V, P, NMAP, NDET = 35, 8, 512, 200, 50
    cfg = make_cfg("ca1m", pst_path=make_pst(P, seed=1), pst_size=P)
    prob = refine_problem(B, V, seed=11)
    Wi, Hi = prob["size"]
    (mt, mR, ms_), (dt, dR, ds) = map_and_detections(NMAP - B, NDET, seed=5)
    far = np.array([40.0, 40.0, 0.0], np.float32)                      # the multi-view boxes live away from the random map
    pt, pR, pp = prob["tensor"].copy(), prob["R"], prob["poses"].copy()
    pt[..., :3] += far; pp[..., :3, 3] += far
    # map = 165 single-view boxes + 35 boxes whose lists hold 8 views (their first view's box stands in the map)
    t_map = np.concatenate([mt, pt[:, 0]]); R_map = np.concatenate([mR, pR[:, 0]]); s_map = np.concatenate([ms_, prob["scores"][:, 0]])
    n1 = NMAP - B
    eye = np.eye(4, dtype=np.float32); eye[:3, 3] = (0, 0, 1.0)
    st_t = np.concatenate([mt, pt.reshape(-1, 6)]); st_R = np.concatenate([mR, pR.reshape(-1, 3, 3)])
    st_s = np.concatenate([ms_, prob["scores"].reshape(-1)]); st_p = np.concatenate([np.tile(eye, (n1, 1, 1)), pp.reshape(-1, 4, 4)])
    M = st_t.shape[0]
    d = str(dev)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(d)       # noqa: E731
    per = api.Instances3D((Hi, Wi))
    per.pred_boxes_3d = api.GeneralInstance3DBoxes(T(st_t), T(st_R))
    per.scores, per.cam_pose = T(st_s), T(st_p)
    per.pred_boxes, per.pred_proj_xy = torch.zeros(M, 4, device=d), torch.zeros(M, 2, device=d)
    per.frame_id, per.init_id, per.valid_num = torch.zeros(M, dtype=torch.int64), torch.arange(M), torch.zeros(M)
    per.project_3d_boxes(prob["K"], H=Hi, W=Wi)
    allp = api.Instances3D((Hi, Wi))
    allp.pred_boxes_3d = api.GeneralInstance3DBoxes(T(t_map), T(R_map))
    first = np.concatenate([np.arange(n1), n1 + V * np.arange(B)])
    allp.scores, allp.cam_pose = T(s_map), T(st_p[first])
    allp.pred_boxes, allp.pred_proj_xy = torch.zeros(NMAP, 4, device=d), torch.zeros(NMAP, 2, device=d)
    allp.frame_id, allp.init_id, allp.valid_num = torch.zeros(NMAP, dtype=torch.int64), torch.from_numpy(first), torch.zeros(NMAP)
    allp.projected_boxes = per.projected_boxes[torch.from_numpy(first).to(d)]
    bm = api.BoxManager(cfg)
    bm.fusion_list = [[i] for i in range(n1)] + [list(range(n1 + V * b, n1 + V * (b + 1))) for b in range(B)]
    bm.fusion_flag = [0] * M
    sess = fastpath.Session(bm, cfg, dev, map_capacity=1024, store_capacity=4096, fused_capacity=1024, max_det=64)
    sess.import_state(allp, per, bm)
    eng = sess.engine
    saved = {"map": {k: v.clone() for k, v in eng.map.items()}, "fflag": eng.fflag.clone(), "fcount": eng.fused["count"].clone()}
    # the keyframe: 50 detections (70 % re-observe map boxes), camera-frame = world frame under an identity-like pose
    pose = eye.copy()
    inv = np.linalg.inv(pose)
    c_cam = (dt[:, :3] - pose[:3, 3]) @ pose[:3, :3]
    tc = np.concatenate([c_cam, dt[:, 3:]], 1).astype(np.float32)
    Rc = np.einsum("ji,njk->nik", pose[:3, :3], dR).astype(np.float32)
    packed = torch.from_numpy(pack_keyframe(tc, Rc, ds, np.zeros((NDET, 4), np.float32), np.zeros((NDET, 2), np.float32), pose, prob["K"], (Wi, Hi), 1)).pin_memory()
    del inv

    def restore():
        for k, v in saved["map"].items():
            eng.map[k].copy_(v)
        eng.fflag.copy_(saved["fflag"]); eng.fused["count"].copy_(saved["fcount"])
        eng._check(eng.lib.bf_engine_set_counts(eng.e, NMAP, M, eng._st()), "bf_engine_set_counts")
        eng.M, eng._n_ub, eng._state_fresh = M, NMAP, False

    med, best = _time_local(lambda: eng.step(packed, NDET), dev, reps=100, warmup=3, restore=restore)
    s = eng.state()
    evals = eng.refine_evals(int(s.B))
    eng.check_status()
    return {"workload": "BASELINE configs[0]: one fusion step, 50 detections vs 200-box map (N=250), 35 boxes x 8 views x 512 particles "
                        "(P=512: 500 is not a multiple of 32, SURVEY H5), one captured CUDA-graph launch incl. the H2D of the detections",
            "ms_median_of_100": round(med, 4), "ms_best": round(best, 4), "launches_per_step": eng.launch_counts[7],
            "refined_boxes": int(s.B), "views": int(s.SV), "evals": evals, "map_after": int(s.N),
            "refine_fp32_frac_upper_bound": round(evals * FLOP_PER_EVAL / (med * 1e-3) / 1e12 / fp32_peak, 4)}


def block_iou(dev, cpu_pairs=20000, cpu=True):
    """BASELINE configs[2]: 256 x 4096 oriented-3D IoU matrix + 3-D NMS over the 4352 boxes."""
    from boxfusion_b200 import ops
    from boxfusion_b200.synthetic import map_and_detections
    (mt, mR, ms_), (dt, dR, ds) = map_and_detections(4096, 256, seed=3, tilt_noise=0.0)
    T = lambda a: torch.from_numpy(a).to(dev)                           # noqa: E731
    ca, cb = ops.box_corners(T(dt), T(dR)), ops.box_corners(T(mt), T(mR))
    fp64_peak = ops.probe_fp64()
    out = {"workload": "BASELINE configs[2]: 256 detections x 4096 map boxes IoU matrix + 3-D NMS over N = 4352",
           "fp64_fma_peak_tflops_measured": round(fp64_peak, 2)}
    for mode, nm in ((ops.IOU_SAMPLED_REF, "sampled_ref"), (ops.IOU_ANALYTIC, "analytic")):
        res = {}
        med, best = _time_local(lambda: res.__setitem__("o", ops.iou3d_matrix(ca, cb, mode=mode, want_stats=True)), dev, reps=20)
        st = res["o"][1].cpu().numpy()
        pairs, gate, ana = int(st[0]), int(st[2]), int(st[3])
        flops = pairs * FLOP_PER_PAIR_GATE + gate * FLOP_PER_PAIR_ESTIMATE if mode == ops.IOU_SAMPLED_REF else \
            ana * FLOP_PER_PAIR_ANALYTIC + gate * FLOP_PER_PAIR_ESTIMATE
        out[nm] = {"ms": round(med, 4), "pairs_per_s": round(pairs / (med * 1e-3), 1), "aabb_pass": int(st[1]), "gate_pass": gate,
                   "analytic_pairs": ana, "gate_pass_fraction": round(gate / pairs, 6),
                   "algorithmic_tflops": round(flops / (med * 1e-3) / 1e12, 3),
                   "frac_of_fp64_peak": round(flops / (med * 1e-3) / 1e12 / fp64_peak, 4),
                   "hbm_bytes_algorithmic": int(4352 * 96 + pairs * 8)}
    t = np.concatenate([mt, dt]); R = np.concatenate([mR, dR]); sc = np.concatenate([ms_, ds])
    n = t.shape[0]
    corners, centers = ops.box_corners(T(t), T(R), want_centers=True)
    scores = T(sc)
    iid = torch.arange(n, dtype=torch.int32, device=dev)
    poses = torch.eye(4, device=dev).reshape(1, 16).repeat(n, 1).contiguous()
    res = {}

    def nms():
        fl = torch.zeros((n, ops.FUSION_CAP), dtype=torch.int32, device=dev); fl[:, 0] = iid
        ln = torch.ones(n, dtype=torch.int32, device=dev); fg = torch.zeros(n, dtype=torch.int32, device=dev)
        order = ops.score_order(scores)                                 # on the device (bf_score_order, counting path for N > 4096)
        res["o"] = ops.nms3d(corners, centers, order, iid, poses, fl, ln, fg, 0.1, 0.8, 30.0, 0.5, ops.IOU_SAMPLED_REF)
    med, best = _time_local(nms, dev, reps=20)
    out["nms_4352"] = {"ms": round(med, 4), "pairs": n * (n - 1) // 2, "pairs_per_s": round(n * (n - 1) / 2 / (med * 1e-3), 1),
                       "kept": int(res["o"][0].sum().item()), "includes": "score order + rank + planes + pairs + gate/counts + greedy matching"}
    if cpu:
        # the reference's own rate: calculate_obb_iou (scipy/Qhull + 25^3 sampling) on uniformly sampled pairs, 1 core
        from oracle import port
        port.IOU_BACKEND = "scipy"
        rs = np.random.RandomState(0)
        ia, ib = rs.randint(0, 256, cpu_pairs), rs.randint(0, 4096, cpu_pairs)
        ca_h, cb_h = ca.cpu().numpy(), cb.cpu().numpy()
        t0 = time.perf_counter()
        done = 0
        for a, b in zip(ia, ib):
            port.Instances3D.obb_iou(ca_h[a], cb_h[b])
            done += 1
            if time.perf_counter() - t0 > 12.0:
                break
        dtc = time.perf_counter() - t0
        out["cpu_reference_port"] = {"pairs_per_s": round(done / dtc, 1), "pairs_sampled": done, "cores": 1,
                                     "note": "oracle/port.py obb_iou (scipy ConvexHull + 25^3 sampling like instances.py:573-613) on uniformly "
                                             "sampled pairs of the same matrix"}
    return out


def block_c4(dev, fp32_peak, reps=5):
    """BASELINE configs[3]: 4096 particles x 32 views x 128 boxes, 20 forced iterations and the early-stop run."""
    from boxfusion_b200 import ops
    from boxfusion_b200.synthetic import make_pst, refine_problem
    B, V, P = 128, 32, 4096
    prob = refine_problem(B, V, seed=11)
    Wi, Hi = prob["size"]
    pst = torch.from_numpy(make_pst(P, seed=1)).to(dev)
    cfg = make_cfg("ca1m", pst_path=None, pst_size=P)
    K16 = np.eye(4, dtype=np.float32); K16[:3, :3] = prob["K"]
    T = lambda a: torch.from_numpy(a).to(dev)                           # noqa: E731
    t, R, s, po = T(prob["tensor"].reshape(-1, 6)), T(prob["R"].reshape(-1, 9)), T(prob["scores"].reshape(-1)), T(prob["poses"].reshape(-1, 16))
    uv = ops.project_boxes(ops.box_corners(t, R), torch.linalg.inv(po.reshape(-1, 4, 4)), prob["K"], Wi, Hi).reshape(-1, 16)
    off = torch.arange(B + 1, dtype=torch.int32, device=dev) * V
    idx = torch.arange(B * V, dtype=torch.int32, device=dev)
    out = {"workload": "BASELINE configs[3]: 4096 particles x 32 views x 128 boxes, one bf_refine launch"}
    for es, nm in ((False, "forced_20_iterations"), (True, "early_stop")):
        rcfg = ops.make_refine_cfg(cfg, K16.reshape(-1), Hi, Wi, early_stop=es)
        res = {}
        med, best = _time_local(lambda: res.__setitem__("o", ops.refine(pst, t, R, s, uv, po, off, idx, rcfg, max_views=V)), dev, reps=reps)
        its = res["o"][2].cpu().numpy()
        evals = float(its.sum()) * P * V
        out[nm] = {"ms": round(med, 3), "evals_per_s": round(evals / (med * 1e-3), 1), "iters_mean": round(float(its.mean()), 2),
                   "ms_per_iteration": round(med / float(its.max()), 4), "launch": ops.last_refine_launch(),
                   "fp32_tflops_algorithmic": round(evals * FLOP_PER_EVAL / (med * 1e-3) / 1e12, 3),
                   "frac_of_fp32_peak": round(evals * FLOP_PER_EVAL / (med * 1e-3) / 1e12 / fp32_peak, 4)}
    return out


def block_c5(args, rank, world, dev):
    """BASELINE configs[4]: 64 independent ScanNet-shaped sequences sharded round-robin over the ranks (strong scaling), every
    rank driving `concurrent` engines at a time on private streams (one C call per keyframe each), no data-path collective."""
    from boxfusion_b200 import ops
    from boxfusion_b200.engine import FusionEngine, pack_keyframe
    from boxfusion_b200.sharding import gather_maps, shard_sequences
    n_seq, frames = args.sequences, args.c5_frames
    cfg = make_cfg("scannet", pst_path=GOLDEN_PST, pst_size=1024)
    mine = shard_sequences(n_seq, rank, world)
    data = []
    for sidx in mine:
        sc = SyntheticScene(n_objects=N_OBJECTS, seed=5000 + sidx, max_det=MAX_DET, shape="scannet")
        kfs = [sc.keyframe(k) for k in range(frames)]
        data.append([(torch.from_numpy(pack_keyframe(k.tensor_cam, k.R_cam, k.scores, k.pred_boxes, k.pred_proj_xy, k.pose, k.K,
                                                     k.image_size, i)).pin_memory(), k.tensor_cam.shape[0]) for i, k in enumerate(kfs)])
    S = max(1, min(args.concurrent, len(data)))
    # with several sequences in flight the refinement uses 256-thread CTAs that share SMs (measured on one B200, 8 / 16 sequences
    # at a time: 13 600 / 17 200 keyframes/s, against 10 500 / 11 400 with the shape that minimises the latency of a lone keyframe)
    shape = args.c5_shape if args.c5_shape != "auto" else ("throughput" if S >= 2 else "latency")
    pool = [FusionEngine(cfg, device=dev, store_capacity=max(32768, 64 * frames), fused_capacity=8192, private_stream=True,
                         concurrent=(shape == "throughput")) for _ in range(S)]

    def run_all(seqs):
        last = None
        for g0 in range(0, len(seqs), S):
            group = seqs[g0:g0 + S]
            engines = pool[: len(group)]
            for e in engines:
                e.reset()
            for k in range(max(len(q) for q in group)):
                for e, q in zip(engines, group):
                    if k < len(q):
                        e.step(q[k][0], q[k][1])
            for e in engines:
                e.check_status()
            last = engines[-1]
        return last

    run_all([data[i % len(data)][: max(args.warmup, 3) + 10] for i in range(S)])      # warm every engine
    ops.Profile.reset()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    last = run_all(data)
    torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) / 1e3
    if world > 1:
        tt = torch.tensor([t], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        t = float(tt[0])
        snap = last.map
        rows = torch.cat([snap["tensor"][: last.N], snap["R"][: last.N]], 1).contiguous()
        assert len(gather_maps(rows)) == world
    return {"workload": f"BASELINE configs[4]: {n_seq} independent synthetic ScanNet-shaped sequences x {frames} keyframes (640x480, 200 objects, "
                        f"<=50 detections/keyframe), sharded round-robin over the ranks, {S} engines at a time per rank on private streams",
            "refine_shape": shape, "keyframes_per_s": round(n_seq * frames / t, 2), "sequences_per_s": round(n_seq / t, 3), "n_gpus": world, "scaling": "strong",
            "wall_s_max_over_ranks": round(t, 3), "gpu_launches_rank0": int(ops.Profile.launches)}


# =====================================================================================================================
def _time_sharded(fn, world, dev, steps, warmup):
    """W warm-up + K timed repetitions of one sharded step, CUDA events, L2 flushed between repetitions, max over ranks."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(max(warmup, 3)):
        fn()
    ms = []
    for _ in range(steps):
        flush.zero_()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        a, b = _events()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    t = torch.tensor([float(np.median(ms))], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t[0])


def run_c4(args, rank, world, local_rank):
    """BASELINE configs[3] with the boxes of the one call sharded over the ranks (boxfusion_b200/sharding.py::refine_sharded,
    SURVEY 8(e) axis 2) + all_gather of the fused rows."""
    from boxfusion_b200 import ops
    from boxfusion_b200.sharding import refine_sharded
    from boxfusion_b200.synthetic import make_pst, refine_problem
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, V, P = 128, 32, 4096
    prob = refine_problem(B, V, seed=11)
    Wi, Hi = prob["size"]
    pst = torch.from_numpy(make_pst(P, seed=1)).to(dev)
    cfg = make_cfg("ca1m", pst_path=None, pst_size=P)
    K16 = np.eye(4, dtype=np.float32); K16[:3, :3] = prob["K"]
    t = torch.from_numpy(prob["tensor"].reshape(-1, 6)).to(dev); R = torch.from_numpy(prob["R"].reshape(-1, 9)).to(dev)
    s = torch.from_numpy(prob["scores"].reshape(-1)).to(dev); po = torch.from_numpy(prob["poses"].reshape(-1, 16)).to(dev)
    uv = ops.project_boxes(ops.box_corners(t, R), torch.linalg.inv(po.reshape(-1, 4, 4)), prob["K"], Wi, Hi).reshape(-1, 16)
    off = np.arange(B + 1, dtype=np.int32) * V
    idx = torch.arange(B * V, dtype=torch.int32, device=dev)
    rcfg = ops.make_refine_cfg(cfg, K16.reshape(-1), Hi, Wi, early_stop=False)
    res = {}
    if world > 1:
        fn = lambda: res.__setitem__("o", refine_sharded(pst, t, R, s, uv, po, off, idx, rcfg))       # noqa: E731
    else:
        fn = lambda: res.__setitem__("o", ops.refine(pst, t, R, s, uv, po, off, idx, rcfg, max_views=V)[:3])   # noqa: E731
    ms = _time_sharded(fn, world, dev, min(args.steps, 10), args.warmup)
    out, upd, its = res["o"]
    assert out.shape[0] == B and int(its.min()) == 20
    if rank != 0:
        return
    evals = float(B) * V * P * 20
    print(json.dumps({"metric": "particle-view evaluations/s, BASELINE configs[3] (4096 x 32 x 128, 20 forced iterations), boxes sharded over the GPUs",
                      "value": round(evals / (ms * 1e-3), 1), "unit": "evals/s", "n_gpus": world, "steps": min(args.steps, 10), "warmup": max(args.warmup, 3),
                      "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": {"workload": "C4 particle sweep, one bf_refine call per rank on its block of boxes + all_gather of [B,6] rows",
                                                      "l2": "flushed between repetitions"},
                      "checksum_updated_boxes": int(upd.sum().item())}))


def run_c3(args, rank, world, local_rank):
    """BASELINE configs[2] with the N = 4352 NMS sharded by rows of the pair triangle over the ranks: every rank finds the
    over-threshold edges of its rows, ONE all_gather of the edge lists (a few KB), the greedy scan runs replicated
    (SURVEY 8(e) axis 3)."""
    from boxfusion_b200 import ops
    from boxfusion_b200.sharding import nms3d_sharded
    from boxfusion_b200.synthetic import map_and_detections
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    (mt, mR, ms_), (dt, dR, ds) = map_and_detections(4096, 256, seed=3, tilt_noise=0.0)
    t = np.concatenate([mt, dt]); R = np.concatenate([mR, dR]); sc = np.concatenate([ms_, ds])
    n = t.shape[0]
    corners, centers = ops.box_corners(torch.from_numpy(t).to(dev), torch.from_numpy(R).to(dev), want_centers=True)
    scores = torch.from_numpy(sc).to(dev)
    iid = torch.arange(n, dtype=torch.int32, device=dev)
    poses = torch.eye(4, device=dev).reshape(1, 16).repeat(n, 1).contiguous()
    res = {}

    def fn():
        fl = torch.zeros((n, ops.FUSION_CAP), dtype=torch.int32, device=dev); fl[:, 0] = iid
        ln = torch.ones(n, dtype=torch.int32, device=dev); fg = torch.zeros(n, dtype=torch.int32, device=dev)
        order = ops.score_order(scores)
        if world > 1:
            res["o"] = nms3d_sharded(corners, centers, order, iid, poses, fl, ln, fg, 0.1, 0.8, 30.0, 0.5, ops.IOU_SAMPLED_REF)
        else:
            res["o"] = ops.nms3d(corners, centers, order, iid, poses, fl, ln, fg, 0.1, 0.8, 30.0, 0.5, ops.IOU_SAMPLED_REF)
    ms = _time_sharded(fn, world, dev, min(args.steps, 50), args.warmup)
    keep = res["o"][0]
    if rank != 0:
        return
    pairs = n * (n - 1) // 2
    print(json.dumps({"metric": "oriented-3D-IoU pairs/s through 3-D NMS, BASELINE configs[2] (N = 4352, SAMPLED_REF), pair rows sharded over the GPUs",
                      "value": round(pairs / (ms * 1e-3), 1), "unit": "pairs/s", "n_gpus": world, "steps": min(args.steps, 50),
                      "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "C3 NMS: every rank finds the over-threshold pairs of its rows, all_gather of the edge lists, greedy scan replicated",
                                 "l2": "flushed between repetitions"},
                      "checksum_kept": int(keep.sum().item())}))


# =====================================================================================================================
# CPU arms
# =====================================================================================================================
def cpu_port_sampled(frames, timed, budget_s, warm_frames=0):
    """Reference algorithm on the host (oracle/port.py): the timed keyframes with the scipy/Qhull IoU + the C restatement of
    the kernel on ONE core (the reference is single-threaded Python); the keyframes in between advance the state with the
    oracle's fast C backend (bit-identical state).  Stops after `budget_s` seconds of timed work."""
    from oracle import port, refine_oracle
    cfg = make_cfg("ca1m", pst_path=GOLDEN_PST, pst_size=1024)
    sess = FusionSession(port, cfg)
    done, spent, ks = 0, 0.0, []
    for k, kf in enumerate(frames):
        if k in timed:
            port.IOU_BACKEND = "scipy"
            refine_oracle.set_threads(1)
            t0 = time.perf_counter()
            sess.step(kf)
            spent += time.perf_counter() - t0
            done += 1
            ks.append(k)
            if spent > budget_s:
                break
        else:
            port.IOU_BACKEND = "c_batch"
            refine_oracle.set_threads(os.cpu_count() or 1)
            sess.step(kf)
    return done, spent, ks, len(sess.all_pred_box)


def cpu_port_openmp(frames, budget_s):
    from oracle import port, refine_oracle
    port.IOU_BACKEND = "c_batch"
    refine_oracle.set_threads(os.cpu_count() or 1)
    cfg = make_cfg("ca1m", pst_path=GOLDEN_PST, pst_size=1024)
    sess = FusionSession(port, cfg)
    t0 = time.perf_counter()
    done = 0
    for kf in frames:
        sess.step(kf)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0, len(sess.all_pred_box)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=FRAMES_PER_SEQUENCE)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--workload", default="c2", choices=["c2", "c5", "c4", "c3"],
                    help="c2 = the bench line (with the other configs as blocks); c5 = only the 64 sharded sequences; c4 / c3 = one large "
                         "step sharded inside (boxes / NMS pair rows)")
    ap.add_argument("--blocks", default="c1,iou,c4,c5", help="extra blocks of the c2 line (comma separated; empty = none)")
    ap.add_argument("--sequences", type=int, default=64)
    ap.add_argument("--c5-frames", type=int, default=FRAMES_PER_SEQUENCE)
    ap.add_argument("--concurrent", type=int, default=16, help="c5: sequences driven concurrently per GPU (streams)")
    ap.add_argument("--c5-shape", default="auto", choices=["auto", "latency", "throughput"], help="c5: refinement launch shape of the engines")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    blocks = [b for b in args.blocks.split(",") if b]

    if args.impl == "reference":
        # the reference's own CPU implementation of the path.  Its association half is Python (numpy/scipy) and cannot
        # travel to the GPU box, so this arm times the oracle port of it (same algorithm and cost structure: Qhull per
        # pair, 25^3 sampling) with the reference's kernel arithmetic in C; single-threaded like the reference.
        if rank != 0:
            return
        K, W = args.steps, max(args.warmup, 3)
        frames, timed = plan(min(K, FRAMES_PER_SEQUENCE), 0)[0]
        warm = build_keyframes(999, W + 12)
        cpu_port_sampled(warm, set(range(W)), 1e9)                      # W warm-up steps of another sequence
        budget = 150.0
        done, dt, ks, nmap = cpu_port_sampled(frames, timed, budget)
        v = done / dt
        print(json.dumps({
            "impl": "reference", "metric": METRIC,
            "value": round(v, 4), "unit": "keyframes/s", "n_gpus": args.gpus, "steps": done, "warmup": W,
            "ms_per_step": round(1e3 * dt / done, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(K),
            "cpu_baseline": {"value": round(v, 4), "unit": "keyframes/s", "cores": 1, "kind": "port",
                             "sample": f"timed keyframes {ks} of the sequence ({done} of {len(timed)} within a {budget:.0f} s budget; map at "
                                       f"{nmap} boxes when it stopped; the later keyframes the budget cut off are the slow ones on the CPU: "
                                       "association is O(N^2) Qhull calls); oracle/port.py with scipy/Qhull IoU, single-threaded like the reference"},
            "e2e": {"value": round(v, 4), "unit": "keyframes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.workload in ("c4", "c3"):
        {"c4": run_c4, "c3": run_c3}[args.workload](args, rank, world, local_rank)
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if args.workload == "c5":
        r = block_c5(args, rank, world, dev)
        if rank == 0:
            print(json.dumps({"metric": "fusion keyframes/s over independent sequences (BASELINE configs[4])", "value": r["keyframes_per_s"],
                              "unit": "keyframes/s", "n_gpus": world, "steps": args.c5_frames, "warmup": max(args.warmup, 3),
                              "ms_per_step": round(1e3 * r["wall_s_max_over_ranks"] * world / (args.sequences * args.c5_frames), 4),
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": r["workload"]}, "sequences_per_s": r["sequences_per_s"],
                              "gpu_launches": r["gpu_launches_rank0"], "wall_s": r["wall_s_max_over_ranks"]}))
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    res = run_ours(args, rank, world, local_rank)
    c5 = block_c5(args, rank, world, dev) if "c5" in blocks else None
    if rank == 0:
        out, seqs, step_ms_e2e = res
        from boxfusion_b200 import ops
        fp32_peak = out["roofline"]["peak"] if out["roofline"] else ops.probe_fp32()
        if c5 is not None:
            out["c5"] = c5
        if world == 1:
            if "c1" in blocks:
                out["c1_step"] = block_c1(dev, fp32_peak)
            if "iou" in blocks:
                out["iou"] = block_iou(dev, cpu=args.cpu_budget > 0)
            if "c4" in blocks:
                out["c4"] = block_c4(dev, fp32_peak)
        if world == 1 and args.cpu_budget > 0:
            frames, timed = seqs[0]
            done, dt, ks, nmap = cpu_port_sampled(frames, timed, args.cpu_budget)
            gpu_same = done / (sum(step_ms_e2e[:done]) / 1e3)
            # a much stronger CPU arm than the reference's Python: the C restatement with OpenMP on every host core
            done_c, dt_c, nmap_c = cpu_port_openmp(frames, max(5.0, args.cpu_budget / 2))
            out["cpu_baseline_c_openmp"] = {
                "value": round(done_c / dt_c, 3), "unit": "keyframes/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"keyframes 0..{done_c - 1} of the sequence, all of them (map grew to {nmap_c} boxes); oracle/*.c with OpenMP over IoU "
                          "pairs and particles - not the reference's implementation, reported to show the gap to an optimised multi-core CPU code"}
            out["cpu_baseline"] = {
                "value": round(done / dt, 4), "unit": "keyframes/s", "cores": 1, "kind": "port",
                "sample": f"the first {done} timed keyframes {ks} of the same sequence within a {args.cpu_budget:.0f} s budget (map at {nmap} boxes); "
                          "single-threaded like the reference; the keyframes in between advanced the state untimed",
                "ours_e2e_on_same_sample": round(gpu_same, 2)}
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    # exactly ONE line on stdout - the JSON: anything a library prints there meanwhile (NCCL's version banner under
    # NCCL_DEBUG=VERSION, for one) goes to stderr instead
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _buf = io.StringIO()
    with contextlib.redirect_stdout(_buf):
        main()
    sys.stdout.flush()
    os.dup2(_real_stdout, 1)
    os.close(_real_stdout)
    sys.stdout.write(_buf.getvalue())
    sys.stdout.flush()
