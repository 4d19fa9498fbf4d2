"""Device-resident fusion engine.

`FusionEngine.step()` is one keyframe of demo.py:200-327 with *all* state resident in HBM: the global map
(`all_pred_box`), the per-frame observation store (`per_frame_ins`) and BoxManager's `fusion_list`,
`fusion_flag`, `already_fusion`.  Round 2: a keyframe is ONE C call (`bf_engine_step`, include/boxfusion_b200.h) that
copies the packed detections to the device and replays one CUDA graph; every size that varies per keyframe lives in
device memory, so there is no read-back and no host decision inside a keyframe.  The host only learns the row
counts when it asks (`N`, `snapshot()`, `export()`, `check_status()`).

It computes exactly what the reference-shaped API (`Instances3D.spatial_association`, `correspondence_association`,
`BoxManager.update`, `BoxFusion.boxfusion`) computes - tests/test_gpu_engine.py compares every field after every
keyframe with the reference goldens - but is not itself part of the reference's interface; `export()` materialises
the reference-shaped containers from the device state.  The reference-shaped API itself runs on this engine when its
containers are used the way demo.py uses them (boxfusion_b200/fastpath.py).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib, ops
from ._lib import (EngineBuffers, EngineCfg, EngineState, FusedTable, KF_HEADER, KF_ROW, MapBuffers, StoreBuffers)
from .box_fusion import _load_pst

_MAP_FIELDS = (("tensor", 6, torch.float32), ("R", 9, torch.float32), ("scores", 1, torch.float32),
               ("box2d", 4, torch.float32), ("projxy", 2, torch.float32), ("pose", 16, torch.float32),
               ("uv", 16, torch.float32), ("valid", 1, torch.float32), ("init_id", 1, torch.int32),
               ("frame_id", 1, torch.int32), ("fl", ops.FUSION_CAP, torch.int32), ("flen", 1, torch.int32))
_STORE_FIELDS = (("tensor", 6), ("R", 9), ("scores", 1), ("uv", 16), ("pose", 16))
STATUS_NAMES = ("nms fusion lists", "correspondence fusion lists", "refine (views / polygon candidates)", "fused table",
                "map / store / detection capacity", "views per box", "IoU work list", "reserved")


def keyframe_header(n: int, frame_id: int, K, image_size, pose) -> np.ndarray:
    """The BF_KF_HEADER words of a packed keyframe.  The inverses are taken on the host with the reference's own
    calls: torch.linalg.inv (instances.py:350) for the observation projection and np.linalg.inv (instances.py:680)
    for the correspondence projection."""
    pose = np.ascontiguousarray(pose, dtype=np.float32).reshape(4, 4)
    hdr = np.empty(KF_HEADER, dtype=np.float32)
    hi = hdr.view(np.int32)
    hi[0], hi[1] = int(n), int(frame_id)
    K = np.asarray(K, dtype=np.float32)
    hdr[2:8] = (K[0, 0], K[1, 1], K[0, 2], K[1, 2], float(image_size[0]), float(image_size[1]))
    hdr[8:24] = pose.reshape(-1)
    hdr[24:40] = torch.linalg.inv(torch.from_numpy(pose)[None])[0].numpy().reshape(-1)
    hdr[40:56] = np.linalg.inv(pose).astype(np.float32).reshape(-1)
    return hdr


def pack_keyframe(tensor_cam, R_cam, scores, pred_boxes, pred_proj_xy, pose, K=None, image_size=None, frame_id: int = 0) -> np.ndarray:
    """One contiguous float32 buffer per keyframe (the engine's only H2D): header (n, frame id, intrinsics, pose and its
    two inverses), then tensor_cam[n,6] R_cam[n,9] scores[n] box2d[n,4] projxy[n,2].  K / image_size may be left out here
    and given to FusionEngine.step instead."""
    n = int(np.asarray(scores).shape[0])
    hdr = keyframe_header(n, frame_id, np.eye(3) if K is None else K, (0, 0) if image_size is None else image_size, pose)
    parts = [np.asarray(a, dtype=np.float32).reshape(-1) for a in (tensor_cam, R_cam, scores, pred_boxes, pred_proj_xy)]
    return np.ascontiguousarray(np.concatenate([hdr] + parts))


class FusionEngine:
    def __init__(self, cfg: dict, device="cuda", map_capacity: int = 4096, store_capacity: int = 65536,
                 fused_capacity: int = 32768, iou_mode: int = ops.IOU_SAMPLED_REF, private_stream: bool = False,
                 max_det: int = 128, use_graph: bool = True, concurrent: Optional[bool] = None):
        """private_stream=True gives the engine its own CUDA stream, so that several engines - independent sequences - can be
        driven concurrently from one host thread (bench.py --workload c5); every engine has its own scratch in any case."""
        self.cfg = cfg
        self.dev = ops._dev(device if str(device) != "cuda" else None)
        self._dev_index = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
        if map_capacity > 65536:
            raise ValueError("map_capacity is limited to 65536 rows")
        self.lib = _lib.load_library()
        self.stream = torch.cuda.Stream(self.dev) if private_stream else None
        self.iou_mode = iou_mode
        self.ncap, self.mcap, self.fcap = int(map_capacity), int(store_capacity), int(fused_capacity)
        self.max_det = int(min(max_det, map_capacity))
        d = self.dev
        with torch.cuda.device(d):
            self._maps = [self._alloc_map() for _ in range(2)]          # [0] = current map, [1] = compaction target
            self.store = {k: torch.zeros((self.mcap, w), dtype=torch.float32, device=d) for k, w in _STORE_FIELDS}
            self.fflag = torch.zeros(self.mcap, dtype=torch.int32, device=d)
            self.fused = {"lists": torch.zeros((self.fcap, ops.FUSION_CAP), dtype=torch.int32, device=d),
                          "len": torch.zeros(self.fcap, dtype=torch.int32, device=d),
                          "hash": torch.zeros(self.fcap, dtype=torch.int64, device=d),
                          "count": torch.zeros(1, dtype=torch.int32, device=d)}
            self.pst = torch.from_numpy(_load_pst(cfg["box_fusion"]["pst_path"])).to(d)
            torch.cuda.current_stream(d).synchronize()           # buffers are zero-filled before the engine's first launch
        bufs = EngineBuffers()
        for i in range(2):
            bufs.map[i] = MapBuffers(*[self._maps[i][k].data_ptr() for k, _, _ in _MAP_FIELDS])
        bufs.store = StoreBuffers(*[self.store[k].data_ptr() for k, _ in _STORE_FIELDS])
        bufs.fusion_flag = self.fflag.data_ptr()
        bufs.fused = FusedTable(self.fused["lists"].data_ptr(), self.fused["len"].data_ptr(), self.fused["hash"].data_ptr(),
                                self.fused["count"].data_ptr(), self.fcap)
        cam = cfg["cam"]
        self.K16 = np.array([[cam["fx"], 0, cam["cx"], 0], [0, cam["fy"], cam["cy"], 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
        self.H, self.W = cam["H"], cam["W"]
        bm, bf = cfg["association"], cfg["box_fusion"]
        ec = EngineCfg()
        ec.map_capacity, ec.store_capacity, ec.max_det, ec.iou_mode = self.ncap, self.mcap, self.max_det, int(iou_mode)
        ec.nms_threshold, ec.small_threshold = float(bf["nms_threshold"]), float(bm["small_threshold"])
        ec.translation_gap, ec.rotation_gap, ec.center_gap = float(bm["translation_gap"]), float(bm["rotation_gap"]), 0.5
        ec.small_size, ec.small_plus = float(np.float32(bf["small_size"])), float(np.float32(bf["small_size"] + 0.1))
        ec.use_fusion, ec.check_valid = int(bool(bf["use"])), int(bool(bf.get("check_valid")))
        ec.gap, ec.use_graph = int(cfg["data"]["gap"]), int(bool(use_graph))
        ec.concurrent = int(private_stream if concurrent is None else bool(concurrent))
        ec.refine = ops.make_refine_cfg(cfg, self.K16.reshape(-1), self.H, self.W)
        ec.pst, ec.P = self.pst.data_ptr(), int(self.pst.shape[0])
        self._ec, self._bufs = ec, bufs
        e = ctypes.c_void_p()
        rc = self.lib.bf_engine_create(self.dev.index, ctypes.byref(ec), ctypes.byref(bufs), ctypes.byref(e))
        if rc != 0:
            msg = self.lib.bf_engine_last_error(e).decode() if e else self.lib.bf_last_error(None).decode()
            if e:
                self.lib.bf_engine_destroy(e)
            raise RuntimeError(f"bf_engine_create failed ({rc}): {msg}")
        self.e = e
        counts = (ctypes.c_int32 * 9)()
        self.lib.bf_engine_launch_counts(e, counts)
        self.launch_counts = list(counts)
        self.full_mask = (_lib.PH_INGEST | _lib.PH_NMS | _lib.PH_CORR | _lib.PH_COMPACT | (_lib.PH_VALID if ec.check_valid else 0) |
                          (_lib.PH_FUSE if ec.use_fusion else 0) | _lib.PH_FINISH)
        self.fuses_finish = bool(ec.use_fusion and use_graph)       # bf_engine_apply closes the keyframe itself in the captured tail
        self._state = EngineState()
        self._state_fresh = True          # host copy of the counters equals the device's
        self.M = 0                        # observations (= box_count = len(per_frame_ins) = len(fusion_flag)); exact on the host
        self._n_ub = 0                    # upper bound of the map rows
        self.count = 0                    # keyframe counter
        self._step, self._step_dev = self.lib.bf_engine_step, self.lib.bf_engine_step_device
        self._stream_ptr = self.stream.cuda_stream if self.stream is not None else None

    def __del__(self):
        e, self.e = getattr(self, "e", None), None
        if e:
            try:
                self.lib.bf_engine_destroy(e)
            except Exception:
                pass

    def _alloc_map(self):
        d = self.dev
        return {k: torch.zeros((self.ncap, w) if w > 1 else (self.ncap,), dtype=dt, device=d) for k, w, dt in _MAP_FIELDS}

    def _st(self) -> int:
        return self._stream_ptr if self._stream_ptr is not None else _lib._raw_stream(self._dev_index)

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({_lib._ERR.get(rc, rc)}): {self.lib.bf_engine_last_error(self.e).decode()}")

    @property
    def map(self):
        return self._maps[0]

    # ---- counters -------------------------------------------------------------------------------------------------
    def state(self) -> EngineState:
        """The device counters (synchronises the engine's stream when they are not known on the host)."""
        if not self._state_fresh:
            self._check(self.lib.bf_engine_read_state(self.e, ctypes.byref(self._state), self._st()), "bf_engine_read_state")
            ops.Profile.d2h_bytes += ctypes.sizeof(EngineState)
            self._state_fresh = True
            self._n_ub = self._state.N
        return self._state

    @property
    def N(self) -> int:
        """map rows (len(all_pred_box))"""
        return int(self.state().N)

    @property
    def last(self) -> dict:
        s = self.state()
        return {"B": int(s.B), "views": int(s.SV)}

    def refine_evals(self, B: int) -> int:
        """(particle, view) evaluations of the last keyframe's refinement: sum over its boxes of iterations x views x
        evaluated particles (reads the engine's per-box iteration counts; measurement passes only)."""
        if B <= 0:
            return 0
        its = np.zeros(B, dtype=np.int32)
        off = np.zeros(B + 1, dtype=np.int32)
        self._check(self.lib.bf_engine_read_i32(self.e, 0, its.ctypes.data, B, self._st()), "bf_engine_read_i32")
        self._check(self.lib.bf_engine_read_i32(self.e, 1, off.ctypes.data, B + 1, self._st()), "bf_engine_read_i32")
        P = int(self.pst.shape[0])
        n_eval = min(32 * (int(self.cfg["box_fusion"]["pst_size"]) // 32), P)
        return int(np.sum(its.astype(np.int64) * np.diff(off).astype(np.int64)) * n_eval)

    def reset(self) -> None:
        """Start a new sequence in the same buffers (and with the same scratch and graphs)."""
        self._check(self.lib.bf_engine_reset(self.e, self._st()), "bf_engine_reset")
        ctypes.memset(ctypes.byref(self._state), 0, ctypes.sizeof(EngineState))
        self._state_fresh = True
        self.M = self.count = self._n_ub = 0

    def update_intrinsics(self, size, K):                     # box_fusion.py:463-466
        self.H, self.W = size[1], size[0]
        self.K16[:3, :3] = np.asarray(K)

    # ---- one keyframe ---------------------------------------------------------------------------------------------
    def step(self, packed, n: int, K=None, image_size=None, frame_id: Optional[int] = None, phases: int = 0) -> None:
        """One keyframe, asynchronous.  `packed`: pack_keyframe() output as a numpy array, a (pinned) CPU tensor or a CUDA
        tensor.  K / image_size / frame_id, when given, overwrite the header of a host `packed` in place (frame_id is the
        FRAME index the reference stores in frame_id and compares with cfg data.gap, demo.py:217, box_manager.py:151-166;
        default: the keyframe counter)."""
        n = int(n)
        if n == 0:                                            # demo.py:206-212
            self.count += 1
            return
        if n > self.max_det:
            raise RuntimeError(f"FusionEngine: {n} detections in one keyframe; max_det is {self.max_det}")
        if self.M + n > self.mcap:
            raise RuntimeError("FusionEngine capacity exceeded (store_capacity)")
        if self._n_ub + n > self.ncap and self.N + n > self.ncap:     # the bound is refreshed from the device before giving up
            raise RuntimeError("FusionEngine capacity exceeded (map_capacity)")
        if isinstance(packed, torch.Tensor) and packed.is_cuda:
            assert K is None and image_size is None and frame_id is None, "a device-resident keyframe carries its own header"
            rc = self._step_dev(self.e, packed.data_ptr(), n, phases, self._st())
        else:
            arr = packed.numpy() if isinstance(packed, torch.Tensor) else packed
            if K is not None or image_size is not None or frame_id is not None or arr.view(np.int32)[0] != n:
                hi = arr.view(np.int32)
                hi[0] = n
                hi[1] = self.count if frame_id is None else int(frame_id)
                if K is not None:
                    K3 = np.asarray(K, dtype=np.float32)
                    arr[2:6] = (K3[0, 0], K3[1, 1], K3[0, 2], K3[1, 2])
                if image_size is not None:
                    arr[6:8] = (float(image_size[0]), float(image_size[1]))
            ops.Profile.h2d_bytes += 4 * (KF_HEADER + KF_ROW * n)
            rc = self._step(self.e, arr.ctypes.data, n, phases, self._st())
        if rc:
            self._check(rc, "bf_engine_step")
        P = ops.Profile
        P.launches += self.launch_counts[7] if phases == 0 else sum(c for i, c in enumerate(self.launch_counts[:7]) if phases >> i & 1)
        P.calls["bf_engine_step"] = P.calls.get("bf_engine_step", 0) + 1
        if phases == 0 or phases & _lib.PH_FINISH:
            self.M += n
            self._n_ub += n
            self.count += 1
        self._state_fresh = False

    # round-1 names of the two halves of a keyframe (the read-back between them is gone: a step is fully asynchronous)
    def step_launch(self, packed, n: int, K=None, image_size=None, frame_id: Optional[int] = None) -> None:
        self.step(packed, n, K, image_size, frame_id)

    def step_finish(self) -> None:
        return None

    def check_status(self):
        s = self.state()
        bad = [STATUS_NAMES[i] for i in range(8) if s.status[i] != 0]
        if bad:
            raise RuntimeError("FusionEngine: capacity error reported by the device: " + ", ".join(bad))

    # -----------------------------------------------------------------------------------------------------------
    def snapshot(self) -> dict:
        """Everything the reference API would have mutated, downloaded (same keys as FusionSession.snapshot)."""
        self.check_status()
        N, mp = self.N, self.map
        flen = mp["flen"][:N].cpu().numpy()
        fl = mp["fl"][:N].cpu().numpy()
        flat = np.concatenate([fl[i, :flen[i]] for i in range(N)]).astype(np.int64) if N else np.zeros(0, np.int64)
        F = int(self.fused["count"].item())
        al, aln = self.fused["lists"][:F].cpu().numpy(), self.fused["len"][:F].cpu().numpy()
        aflat = np.concatenate([al[i, :aln[i]] for i in range(F)]).astype(np.int64) if F else np.zeros(0, np.int64)
        return {"tensor": mp["tensor"][:N].cpu().numpy(), "R": mp["R"][:N].reshape(N, 3, 3).cpu().numpy(),
                "scores": mp["scores"][:N].cpu().numpy(), "valid_num": mp["valid"][:N].cpu().numpy(),
                "init_id": mp["init_id"][:N].cpu().numpy().astype(np.int64),
                "fusion_flat": flat, "fusion_off": np.cumsum(np.concatenate([[0], flen])).astype(np.int64),
                "fusion_flag": self.fflag[: self.M].cpu().numpy().astype(np.int64),
                "already_flat": aflat, "already_off": np.cumsum(np.concatenate([[0], aln])).astype(np.int64)}

    def export(self):
        """(all_pred_box, per_frame_ins, box_manager) in the reference's container types, built from the device state."""
        from . import api
        self.check_status()
        N, M, mp = self.N, self.M, self.map
        a = api.Instances3D((int(self.H), int(self.W)))
        a.scores, a.pred_boxes, a.pred_proj_xy = mp["scores"][:N].clone(), mp["box2d"][:N].clone(), mp["projxy"][:N].clone()
        a.pred_boxes_3d = api.GeneralInstance3DBoxes(mp["tensor"][:N], mp["R"][:N].reshape(N, 3, 3))
        a.cam_pose = mp["pose"][:N].reshape(N, 4, 4).clone()
        a.frame_id, a.init_id = mp["frame_id"][:N].to(torch.int64), mp["init_id"][:N].to(torch.int64)
        a.valid_num, a.projected_boxes = mp["valid"][:N].clone(), mp["uv"][:N].reshape(N, 8, 2).clone()
        p = api.Instances3D((int(self.H), int(self.W)))
        p.scores = self.store["scores"][:M, 0].clone()
        p.pred_boxes_3d = api.GeneralInstance3DBoxes(self.store["tensor"][:M], self.store["R"][:M].reshape(M, 3, 3))
        p.cam_pose = self.store["pose"][:M].reshape(M, 4, 4).clone()
        p.projected_boxes = self.store["uv"][:M].reshape(M, 8, 2).clone()
        bm = api.BoxManager(self.cfg)
        snap = self.snapshot()
        off, flat = snap["fusion_off"], snap["fusion_flat"]
        bm.fusion_list = [[int(x) for x in flat[off[i]:off[i + 1]]] for i in range(N)]
        bm.fusion_flag = [int(x) for x in snap["fusion_flag"]]
        bm.last_fusion_frame = [[0] for _ in range(M)]
        aoff, aflat = snap["already_off"], snap["already_flat"]
        bm.already_fusion = [[int(x) for x in aflat[aoff[i]:aoff[i + 1]]] for i in range(len(aoff) - 1)]
        return a, p, bm
