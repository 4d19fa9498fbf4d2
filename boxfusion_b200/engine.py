"""Device-resident fusion engine.

`FusionEngine.step()` is one keyframe of demo.py:200-327 with *all* state resident in HBM: the global map
(`all_pred_box`), the per-frame observation store (`per_frame_ins`) and BoxManager's `fusion_list`,
`fusion_flag`, `already_fusion`.  Per keyframe the host issues one H2D copy (the packed detections), a fixed
sequence of library calls, and reads back 32 bytes (row counts for the next launch configuration); results are
downloaded only when asked for (`snapshot()`, `export()`).

It computes exactly what the reference-shaped API (`Instances3D.spatial_association`, `correspondence_association`,
`BoxManager.update`, `BoxFusion.boxfusion`) computes - tests/test_gpu_engine.py compares every field after every
keyframe with the reference goldens - but is not itself part of the reference's interface; `export()` materialises
the reference-shaped containers from the device state.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import ops
from ._lib import FusedTable, Handle, MapBuffers, StoreBuffers, handle, ptr
from .box_fusion import _load_pst

_MAP_FIELDS = (("tensor", 6, torch.float32), ("R", 9, torch.float32), ("scores", 1, torch.float32),
               ("box2d", 4, torch.float32), ("projxy", 2, torch.float32), ("pose", 16, torch.float32),
               ("uv", 16, torch.float32), ("valid", 1, torch.float32), ("init_id", 1, torch.int32),
               ("frame_id", 1, torch.int32), ("fl", ops.FUSION_CAP, torch.int32), ("flen", 1, torch.int32))
_STORE_FIELDS = (("tensor", 6), ("R", 9), ("scores", 1), ("uv", 16), ("pose", 16))


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def pack_keyframe(tensor_cam, R_cam, scores, pred_boxes, pred_proj_xy, pose) -> np.ndarray:
    """One contiguous float32 buffer per keyframe (the engine's only H2D): fields, then pose, then its inverses.
    The inverses are taken on the host with the reference's own calls: torch.linalg.inv (instances.py:350) for the
    observation projection and np.linalg.inv (instances.py:680) for the correspondence projection."""
    pose = np.ascontiguousarray(pose, dtype=np.float32).reshape(4, 4)
    inv_t = torch.linalg.inv(torch.from_numpy(pose)[None])[0].numpy()
    inv_n = np.linalg.inv(pose).astype(np.float32)
    parts = [np.asarray(a, dtype=np.float32).reshape(-1) for a in (tensor_cam, R_cam, scores, pred_boxes, pred_proj_xy)]
    return np.ascontiguousarray(np.concatenate(parts + [pose.reshape(-1), inv_t.reshape(-1), inv_n.reshape(-1)]))


class FusionEngine:
    def __init__(self, cfg: dict, device="cuda", map_capacity: int = 4096, store_capacity: int = 65536,
                 fused_capacity: int = 32768, iou_mode: int = ops.IOU_SAMPLED_REF, private_stream: bool = False):
        """private_stream=True gives the engine its own CUDA stream and its own library handle (scratch), so that
        several engines - independent sequences - can be driven concurrently from one host thread with
        step_launch() / step_finish() (bench.py --workload c5)."""
        self.cfg = cfg
        self.dev = ops._dev(device if str(device) != "cuda" else None)
        if map_capacity > 65536:
            raise ValueError("map_capacity is limited to 65536 rows")
        if private_stream:
            self.h = Handle(self.dev.index if self.dev.index is not None else torch.cuda.current_device())
            self.h.check(self.h.lib.bf_set_option(self.h.h, 1, 1), "bf_set_option")     # BF_OPT_REFINE_CONCURRENT
            self.stream = torch.cuda.Stream(self.dev)
        else:
            self.h = handle(self.dev)
            self.stream = None
        self._evt = torch.cuda.Event()
        self._pending = None
        self.iou_mode = iou_mode
        self.ncap, self.mcap, self.fcap = int(map_capacity), int(store_capacity), int(fused_capacity)
        d = self.dev
        self._maps = [self._alloc_map() for _ in range(2)]          # ping-pong for compaction
        self._cur = 0
        self.store = {k: torch.zeros((self.mcap, w), dtype=torch.float32, device=d) for k, w in _STORE_FIELDS}
        self._store_c = StoreBuffers(*[self.store[k].data_ptr() for k, _ in _STORE_FIELDS])
        self.fflag = torch.zeros(self.mcap, dtype=torch.int32, device=d)
        self.fused = {"lists": torch.zeros((self.fcap, ops.FUSION_CAP), dtype=torch.int32, device=d),
                      "len": torch.zeros(self.fcap, dtype=torch.int32, device=d),
                      "hash": torch.zeros(self.fcap, dtype=torch.int64, device=d),
                      "count": torch.zeros(1, dtype=torch.int32, device=d)}
        self._fused_c = FusedTable(self.fused["lists"].data_ptr(), self.fused["len"].data_ptr(), self.fused["hash"].data_ptr(),
                                   self.fused["count"].data_ptr(), self.fcap)
        self.keep = torch.zeros(self.ncap, dtype=torch.int32, device=d)
        self.success = torch.zeros(self.ncap, dtype=torch.int32, device=d)
        self.todo = torch.zeros(self.ncap, dtype=torch.int32, device=d)
        self.offsets = torch.zeros(self.ncap + 1, dtype=torch.int32, device=d)
        self.view_index = torch.zeros(self.ncap * ops.FUSION_CAP, dtype=torch.int32, device=d)
        self.corners = torch.zeros((self.ncap, 8, 3), dtype=torch.float32, device=d)
        self.centers = torch.zeros((self.ncap, 3), dtype=torch.float32, device=d)
        self.out = torch.zeros((self.ncap, 6), dtype=torch.float32, device=d)
        self.upd = torch.zeros(self.ncap, dtype=torch.int32, device=d)
        self.its = torch.zeros(self.ncap, dtype=torch.int32, device=d)
        self.info = torch.zeros(8, dtype=torch.int32, device=d)
        self.status = torch.zeros(4, dtype=torch.int32, device=d)
        self._info_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.pst = torch.from_numpy(_load_pst(cfg["box_fusion"]["pst_path"])).to(d)
        cam = cfg["cam"]
        self.K16 = np.array([[cam["fx"], 0, cam["cx"], 0], [0, cam["fy"], cam["cy"], 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
        self.H, self.W = cam["H"], cam["W"]
        self.N = 0            # map rows
        self.M = 0            # observations (= box_count = len(per_frame_ins) = len(fusion_flag))
        self.count = 0        # keyframe counter
        self.last = {"B": 0, "views": 0}
        self.refine_log = None
        # every buffer above lives as long as the engine: resolve the device pointers once (42 data_ptr() calls per keyframe
        # otherwise - the engine's host time is what bounds bench.py --workload c5)
        self._p = {k: getattr(self, k).data_ptr() for k in ("corners", "centers", "keep", "success", "todo", "offsets", "view_index",
                                                           "out", "upd", "its", "info", "status", "fflag", "pst")}
        self._p.update({"store_" + k: self.store[k].data_ptr() for k, _ in _STORE_FIELDS})
        for m in self._maps:
            m["_p"] = {k: m[k].data_ptr() for k, _, _ in _MAP_FIELDS}
        self._order = torch.zeros(self.ncap, dtype=torch.int32, device=d)
        self._arange = torch.arange(self.ncap, dtype=torch.int32, device=d)
        self._p["order"] = self._order.data_ptr()
        self._stream_ptr = self.stream.cuda_stream if self.stream is not None else None
        if self.stream is not None:
            torch.cuda.current_stream(self.dev).synchronize()   # buffers were zero-filled on the creating stream

    def _call(self, name, fn, *args):
        """ops._call without the per-call event plumbing (launch accounting kept; timing mode falls back to ops._call)."""
        P = ops.Profile
        if P.timing:
            return ops._call(self.h, name, fn, *args)
        P.launches += ops.KERNELS_PER_CALL[name]
        P.calls[name] = P.calls.get(name, 0) + 1
        rc = fn(*args)
        if rc:
            self.h.check(rc, name)

    def reset(self) -> None:
        """Start a new sequence in the same buffers (and with the same library handle / scratch)."""
        assert self._pending is None
        with self._ctx():
            self.fused["count"].zero_()
            self.status.zero_()
            self.info.zero_()
        self.N = self.M = self.count = 0
        self.last = {"B": 0, "views": 0}

    def _alloc_map(self):
        d = self.dev
        t = {k: torch.zeros((self.ncap, w) if w > 1 else (self.ncap,), dtype=dt, device=d) for k, w, dt in _MAP_FIELDS}
        t["_c"] = MapBuffers(*[t[k].data_ptr() for k, _, _ in _MAP_FIELDS])
        return t

    @property
    def map(self):
        return self._maps[self._cur]

    def update_intrinsics(self, size, K):                     # box_fusion.py:463-466
        self.H, self.W = size[1], size[0]
        self.K16[:3, :3] = np.asarray(K)

    # -----------------------------------------------------------------------------------------------------------
    def step(self, packed, n: int, K, image_size) -> None:
        """One keyframe.  `packed`: pack_keyframe() output as a (pinned) CPU tensor, numpy array or CUDA tensor."""
        self.step_launch(packed, n, K, image_size)
        self.step_finish()

    def _ctx(self):
        return torch.cuda.stream(self.stream) if self.stream is not None else _NullCtx()

    def step_launch(self, packed, n: int, K, image_size) -> None:
        """Issue everything of the keyframe up to the 32-byte read-back (asynchronous)."""
        assert self._pending is None, "step_finish() of the previous keyframe has not been called"
        self.update_intrinsics(image_size, K)                 # demo.py:117-118 (update_K_flag stays False)
        if n == 0:                                            # demo.py:206-212
            self.count += 1
            return
        if self.N + n > self.ncap or self.M + n > self.mcap:
            raise RuntimeError("FusionEngine capacity exceeded (map_capacity / store_capacity)")
        with self._ctx():
            self._launch(packed, n, K, image_size)

    def _launch(self, packed, n, K, image_size):
        h, lib, P = self.h, self.h.lib, self._p
        st = self._stream_ptr if self._stream_ptr is not None else h.stream()
        if isinstance(packed, torch.Tensor):
            if not packed.is_cuda:
                ops.Profile.h2d_bytes += packed.numel() * 4
            buf = packed.to(self.dev, non_blocking=True)
        else:
            buf = ops.dev_tensor(packed, torch.float32, self.dev)
        bufp = buf.data_ptr()
        K3 = np.asarray(K, dtype=np.float32)
        fx, fy, cx, cy = float(K3[0, 0]), float(K3[1, 1]), float(K3[0, 2]), float(K3[1, 2])
        Wf, Hf = float(image_size[0]), float(image_size[1])
        mp = self.map
        mpp = mp["_p"]
        N0, M0 = self.N, self.M
        call = self._call
        call("bf_engine_ingest", lib.bf_engine_ingest, h.h, bufp, n, fx, fy, cx, cy, Wf, Hf, self.count, M0, N0, M0, M0,
             ctypes.byref(mp["_c"]), ctypes.byref(self._store_c), P["fflag"], st)
        self.M = M0 + n
        if N0 == 0:                                           # first keyframe: demo.py:228-243
            self.N = n
            self.count += 1
            return
        Nall = N0 + n
        bm, bf = self.cfg["association"], self.cfg["box_fusion"]
        # STEP 1: spatial association (demo.py:262)
        call("bf_box_corners", lib.bf_box_corners, h.h, mpp["tensor"], mpp["R"], Nall, P["corners"], P["centers"], st)
        if Nall <= ops.ORDER_MAX:                             # scores.argsort()[::-1] (instances.py:52), stable, on the device
            call("bf_score_order", lib.bf_score_order, h.h, mpp["scores"], Nall, P["order"], st)
            orderp = P["order"]
        else:
            order = torch.argsort(mp["scores"][:Nall], descending=True, stable=True).to(torch.int32)
            orderp = order.data_ptr()
        call("bf_nms3d", lib.bf_nms3d, h.h, P["corners"], P["centers"], Nall, orderp, mpp["init_id"],
             P["store_pose"], self.M, mpp["fl"], mpp["flen"], P["fflag"], float(bf["nms_threshold"]),
             float(bm["translation_gap"]), float(bm["rotation_gap"]), 0.5, int(self.iou_mode), P["keep"],
             P["success"], P["status"], st)
        # STEP 2: correspondence association for small objects (demo.py:273-289) + valid_num of STEP 1
        pinv = bufp + 4 * (22 * n + 32)
        small = float(np.float32(bf["small_size"]))
        call("bf_engine_corr", lib.bf_engine_corr, h.h, ctypes.byref(mp["_c"]), P["store_pose"], P["fflag"],
             N0, n, P["keep"], P["success"], pinv, fx, fy, cx, cy, Wf, Hf, small,
             float(np.float32(bf["small_size"] + 0.1)), float(bm["small_threshold"]), float(bm["translation_gap"]),
             float(bm["rotation_gap"]), P["info"], P["status"] + 4, st)
        # all_pred_box[keep_idx]; box_manager.update(keep_idx) (demo.py:292 / 325-327)
        other = self._maps[1 - self._cur]
        call("bf_engine_compact", lib.bf_engine_compact, h.h, P["keep"], Nall, ctypes.byref(mp["_c"]),
             ctypes.byref(other["_c"]), P["info"], st)
        self._cur = 1 - self._cur
        mp = self.map
        if bf.get("check_valid"):
            # BoxManager.check_valid_num (box_manager.py:151-166, demo.py:297-298; only when a new box survived, demo.py:269):
            # drop map rows never re-observed (valid_num == 0) that are older than `gap` keyframes - one more compaction.
            # The row count of the compacted map is still on the device (info[1]), so the flags are formed there.
            thr = self.count - int(self.cfg["data"]["gap"])
            stale = (mp["valid"][:Nall] == 0) & (mp["frame_id"][:Nall] < thr) & (self.info[0] != 0)
            self.keep[:Nall].copy_(((self._arange[:Nall] < self.info[1]) & ~stale).to(torch.int32))
            other = self._maps[1 - self._cur]
            call("bf_engine_compact", lib.bf_engine_compact, h.h, P["keep"], Nall, ctypes.byref(mp["_c"]),
                 ctypes.byref(other["_c"]), P["info"], st)
            self._cur = 1 - self._cur
            mp = self.map
        # STEP 3: multi-view box fusion (demo.py:304-305): selection now, refinement after the read-back
        if bf["use"]:
            call("bf_engine_select", lib.bf_engine_select, h.h, ctypes.byref(mp["_c"]), ctypes.byref(self._fused_c),
                 P["info"], P["todo"], P["offsets"], P["view_index"], st)
        self._info_host.copy_(self.info, non_blocking=True)   # the step's only D2H: 32 bytes
        self._evt.record(torch.cuda.current_stream(self.dev))
        self._pending = True
        self._keep_alive = buf                                # the packed detections until the next keyframe is issued

    def step_finish(self) -> None:
        """Wait for the read-back, then launch the refinement of the selected boxes and its write-back."""
        if self._pending is None:
            return
        self._pending = None
        self._evt.synchronize()
        ops.Profile.d2h_bytes += 32
        bf = self.cfg["box_fusion"]
        info = self._info_host.numpy()
        self.N = int(info[1])
        B, SV, maxV = (int(info[2]), int(info[3]), int(info[4])) if bf["use"] else (0, 0, 0)
        if info[5] != 0:
            raise RuntimeError(f"FusionEngine: a fusion list has {maxV} views; bf_refine supports {ops.MAX_VIEWS}")
        self.last = {"B": B, "views": SV}
        if B > 0:                                             # explicit stream argument: no torch stream context needed here
            h, lib, mp, P = self.h, self.h.lib, self.map, self._p
            st = self._stream_ptr if self._stream_ptr is not None else h.stream()
            rcfg = self._rcfg_cached()
            rcfg.views_total, rcfg.max_views = SV, maxV
            self._call("bf_refine", lib.bf_refine, h.h, P["pst"], self.pst.shape[0], P["store_tensor"],
                       P["store_R"], P["store_scores"], P["store_uv"], P["store_pose"], self.M,
                       P["offsets"], P["view_index"], B, ctypes.byref(rcfg), P["out"], P["upd"], P["its"],
                       None, P["status"] + 8, st)
            self._call("bf_engine_apply", lib.bf_engine_apply, h.h, ctypes.byref(mp["_c"]), ctypes.byref(self._fused_c),
                       P["fflag"], P["info"], P["todo"], P["out"], P["upd"], P["status"] + 12, st)
            if self.refine_log is not None:
                self.refine_log.append((B, SV))
        self.count += 1

    def _rcfg_cached(self):
        """bf_refine_cfg of the current intrinsics (rebuilt only when update_intrinsics changed them)."""
        key = (self.H, self.W, self.K16.tobytes())
        if getattr(self, "_rcfg_key", None) != key:
            self._rcfg_key, self._rcfg = key, ops.make_refine_cfg(self.cfg, self.K16.reshape(-1), self.H, self.W)
        return self._rcfg

    def check_status(self):
        if self.stream is not None:
            self.stream.synchronize()
        s = self.status.cpu().numpy()
        if (s != 0).any():
            raise RuntimeError(f"FusionEngine: capacity error reported by the device (status {s.tolist()})")

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
