"""The reference-shaped API on the device-resident engine (round 2).

demo.py:243-327 makes five calls per keyframe into the hot path - `Instances3D.cat` (twice), `spatial_association`,
`correspondence_association`, `BoxManager.update` [`check_valid_num`], `BoxFusion.boxfusion` - and keeps the state between
them in Python containers.  Round 1 implemented every call on its own: upload what the call needs, launch, download what it
returns, re-index the containers with torch ops (1.6 ms per keyframe, 0.46 ms of it on the GPU).  Here the containers the
calls hand back are *views of the engine's HBM state* that materialise only when somebody looks at them, and every call
replays the matching phase of the engine's captured keyframe (`bf_engine_step(.., phases)`):

    Instances3D.cat([all_pred_box, pred_instances])     -> bf_engine_ingest_world (row copies, no torch.cat)
    Instances3D.spatial_association(...)                -> PH_NMS,  read keep / success flags   (the call returns them)
    Instances3D.correspondence_association(...)         -> PH_CORR, read keep flags, PH_COMPACT (the call returns keep_idx)
    BoxManager.update / check_valid_num                 -> nothing / PH_VALID
    BoxFusion.boxfusion(...)                            -> PH_FUSE + PH_FINISH, no read-back

A `Session` hangs off the BoxManager.  It is entered at the end of an ordinary (call-by-call) `boxfusion()` by importing the
containers into an engine, and it is left - the engine state exported back into plain containers, nothing lost - the moment
a call does not match what demo.py does (other thresholds, edited fusion lists, containers the session did not hand out,
CPU tensors ...).  The call-by-call implementation then carries on, and the next `boxfusion()` imports again.  There is no
CPU path involved either way.

Containers handed out by the session alias the engine's buffers: in-place edits go to the engine, and a container of a
previous keyframe shows the current map once a later keyframe has run (the reference's containers are independent copies;
demo.py never looks back at them).
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Optional

import numpy as np
import torch

from . import _lib, ops
from ._lib import KF_HEADER

ENABLED = True                      # module switch (tests compare the two implementations)
RUN_AHEAD = True                    # queue the rest of the keyframe behind spatial_association (rolled back if the caller strays)
ENGINE_KWARGS: dict = {}       # extra FusionEngine arguments of new sessions (tests: use_graph=False runs the same launches eagerly)

_MAP_NATIVE = ("scores", "pred_boxes", "pred_proj_xy", "pred_boxes_3d", "cam_pose", "frame_id", "init_id", "valid_num", "projected_boxes")
_STORE_NATIVE = ("scores", "pred_boxes_3d", "cam_pose", "projected_boxes")


def _cat_values(a, b):
    if isinstance(a, torch.Tensor):
        return torch.cat([a, b.to(a.device) if isinstance(b, torch.Tensor) else torch.as_tensor(b, device=a.device)], dim=0)
    if isinstance(a, np.ndarray):
        return np.concatenate([a, np.asarray(b)], axis=0)
    if isinstance(a, list):
        return list(a) + list(b)
    if hasattr(type(a), "cat"):
        return type(a).cat([a, b])
    raise ValueError("Unsupported type {} for concatenation".format(type(a)))


def _index_values(v, idx: np.ndarray):
    if isinstance(v, torch.Tensor):
        return v[torch.from_numpy(idx).to(v.device)]
    if isinstance(v, np.ndarray):
        return v[idx]
    if isinstance(v, list):
        return [v[int(i)] for i in idx]
    return v[idx]


class EngineFields(dict):
    """`Instances3D._fields` whose values are views of the engine's buffers, each created when it is first asked for."""

    def __init__(self, sess: "Session", kind: str, rows: Optional[int], names):
        super().__init__()
        self.sess, self.kind, self.rows = sess, kind, rows
        self.pending = set(names)                  # fields not materialised yet

    @property
    def filled(self) -> bool:
        return not self.pending

    def _make(self, k):
        self.pending.discard(k)
        v = self.sess.view(self, k)
        dict.__setitem__(self, k, v)
        return v

    def fill(self):
        for k in list(self.pending):
            self._make(k)

    def n_rows(self) -> int:
        if self.rows is None:
            self.sess.settle()
            self.rows = self.sess.engine.N
        return self.rows

    def __getitem__(self, k):
        if k in self.pending:
            return self._make(k)
        return dict.__getitem__(self, k)

    def __setitem__(self, k, v):
        self.pending.discard(k); dict.__setitem__(self, k, v)

    def __delitem__(self, k):
        if k in self.pending:
            self.pending.discard(k)
        else:
            dict.__delitem__(self, k)

    def __contains__(self, k):
        return k in self.pending or dict.__contains__(self, k)

    def __iter__(self):
        self.fill(); return dict.__iter__(self)

    def __len__(self):
        return len(self.pending) + dict.__len__(self)

    def keys(self):
        self.fill(); return dict.keys(self)

    def values(self):
        self.fill(); return dict.values(self)

    def items(self):
        self.fill(); return dict.items(self)

    def get(self, k, default=None):
        return self[k] if k in self else default

    def pop(self, k, *a):
        if k in self.pending:
            self.pending.discard(k)
            return self.sess.view(self, k)
        return dict.pop(self, k, *a)


def cfg_key(cfg) -> tuple:
    """Every cfg value the engine froze into its captured keyframe (a caller may edit the dict between keyframes)."""
    bf, a = cfg["box_fusion"], cfg["association"]
    ro = bf["random_opt"]
    return (float(bf["nms_threshold"]), float(a["small_threshold"]), float(bf["small_size"]), float(a["translation_gap"]),
            float(a["rotation_gap"]), bool(bf["use"]), bool(bf.get("check_valid")), int(cfg["data"]["gap"]), int(bf["iters"]),
            int(bf["pst_size"]), float(ro["center_init_size"]), float(ro["center_scaling_coefficient"]),
            float(ro["shape_init_size"]), float(ro["shape_scaling_coefficient"]))


def _plain_fields(ins) -> dict:
    f = ins._fields
    if isinstance(f, EngineFields):
        f.fill()
    return f


def list_hash(lst) -> int:
    """e_list_hash of csrc/bf_engine.cu (64-bit FNV-1a over the entries, seeded with the length)."""
    h = 1469598103934665603 ^ len(lst)
    for v in lst:
        h ^= int(v) & 0xffffffff
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


class Session:
    """One sequence's engine + where the keyframe in flight stands in demo.py's call order."""
    IDLE, CAT, NMS, CORR = range(4)

    def __init__(self, bm, cfg, device, map_capacity=4096, store_capacity=65536, fused_capacity=32768, max_det=256):
        from .engine import FusionEngine
        from . import instances as inst_mod
        self.cfg = cfg
        self.key = cfg_key(cfg)
        self.use_fusion, self.check_valid, self.gap = bool(cfg["box_fusion"]["use"]), bool(cfg["box_fusion"].get("check_valid")), int(cfg["data"]["gap"])
        self.map_dev, self.store_dev = {}, {}        # device of every engine-held field in the caller's own containers
        self.engine = FusionEngine(cfg, device=device, map_capacity=map_capacity, store_capacity=store_capacity,
                                   fused_capacity=fused_capacity, max_det=max_det, iou_mode=inst_mod.IOU_MODE, **ENGINE_KWARGS)
        self.iou_mode = inst_mod.IOU_MODE
        self.dev = self.engine.dev
        self.bm = weakref.ref(bm)
        self.stage = Session.IDLE
        self.map_c = None               # the Instances3D that currently stands for all_pred_box
        self.store_c = None             # ... for per_frame_ins
        self.cat_c = None               # ... for cat([all_pred_box, pred_instances]) of the keyframe in flight
        self.n = 0
        self.N = 0                      # map rows known on the host (None after check_valid_num until somebody asks)
        self.M = 0
        self.pred = None                # detections of the keyframe in flight
        self.hdr = None
        self.last_keep = self.last_success = None
        self.map_extras = {}            # fields of all_pred_box the engine does not hold (categories, features ...): name -> value [N]
        self.cat_extras = {}
        self.store_chunks = {}          # fields of per_frame_ins the engine does not hold: name -> list of per-keyframe values
        self.lists_host = True          # BoxManager's Python lists are current
        self.lists_given = None         # copies of the lists handed to the caller while the device is ahead
        self.pending_new = None         # init_new_predictions(n, M) the engine has not seen yet
        self.ahead = False              # the engine has run the whole keyframe; the caller is still at `stage` (run-ahead)
        self.valid_called = False       # the caller has called check_valid_num for the keyframe in flight
        self.any_new = 1
        self.frame_id = -1
        self._keep = np.zeros(self.engine.ncap, dtype=np.int32)
        self._succ = np.zeros(self.engine.ncap, dtype=np.int32)
        self._keep_p, self._succ_p = self._keep.ctypes.data, self._succ.ctypes.data
        self._state_ref = ctypes.byref(self.engine._state)
        self._stp = self.engine._st()   # stream of the keyframe in flight (looked up once per keyframe)
        self.image_size = None
        self._flag_views = {}           # numpy / ctypes views of the engine's pinned flag buffers, per slot
        self._ra_launches = None        # kernels of one run-ahead (constant per engine)

    # ---- helpers ----------------------------------------------------------------------------------------------
    def _phase(self, phases):
        e = self.engine
        rc = e.lib.bf_engine_step(e.e, None, self.n, phases, self._stp)
        if rc:
            e._check(rc, "bf_engine_step")
        ops.Profile.launches += sum(c for i, c in enumerate(e.launch_counts[:7]) if phases >> i & 1)
        e._state_fresh = False
        self.lists_host = False                                    # fusion lists / flags / already_fusion move on the device

    def _read_flags(self, count, want_success):
        e = self.engine
        rc = e.lib.bf_engine_read_flags(e.e, self._keep_p, self._succ_p if want_success else None, count, self._state_ref, self._stp)
        if rc:
            e._check(rc, "bf_engine_read_flags")
        e._state_fresh = True
        ops.Profile.d2h_bytes += 4 * count * (2 if want_success else 1) + ctypes.sizeof(_lib.EngineState)
        st = e._state
        if any(st.status[i] for i in range(8)):
            e.check_status()

    # ---- run-ahead: the engine finishes the keyframe while the caller is still between its calls -----------------------------
    def _run_ahead(self, rows):
        e = self.engine
        rc = e.lib.bf_engine_run_ahead(e.e, rows, self._stp)
        if rc:
            e._check(rc, "bf_engine_run_ahead")
        if self._ra_launches is None:                              # NMS + publish, snapshot + correspondence + publish, the rest
            lc = e.launch_counts
            self._ra_launches = lc[1] + 3 + lc[2] + sum(c for i, c in enumerate(lc[:7]) if i >= 3 and (e.full_mask >> i) & 1) - (1 if e.fuses_finish else 0)
        ops.Profile.launches += self._ra_launches
        e._state_fresh = False
        self.lists_host = False
        self.ahead = True

    def _wait_flags(self, slot, rows, want_success):
        e = self.engine
        pk, ps, pst = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        rc = e.lib.bf_engine_wait_flags(e.e, slot, ctypes.byref(pk), ctypes.byref(ps), ctypes.byref(pst))
        if rc:
            e._check(rc, "bf_engine_wait_flags")
        key = (slot, pk.value, ps.value, pst.value)
        views = self._flag_views.get(key)
        if views is None:                                          # the pinned flag buffers of a slot never move: wrap them once
            I32 = ctypes.POINTER(ctypes.c_int32)
            cap = int(e.ncap)
            views = (np.ctypeslib.as_array(ctypes.cast(pk, I32), shape=(cap,)), np.ctypeslib.as_array(ctypes.cast(ps, I32), shape=(cap,)),
                     ctypes.cast(pst, ctypes.POINTER(_lib.EngineState)).contents)
            self._flag_views[key] = views
        keep = views[0][:rows]
        succ = views[1][:rows] if want_success else None
        st = views[2]
        ops.Profile.d2h_bytes += 4 * rows * (2 if want_success else 1) + ctypes.sizeof(_lib.EngineState)
        s8 = st.status
        if s8[0] | s8[1] | s8[2] | s8[3] | s8[4] | s8[5] | s8[6] | s8[7]:
            e.check_status()
        return keep, succ, st

    def settle(self):
        """Somebody is about to look at engine state (or the caller strayed) while the engine has run ahead of the caller:
        put the engine where the caller is - restore the snapshot taken right after the NMS phase and re-issue what the
        caller has called since."""
        if not self.ahead:
            return
        self.ahead = False
        if self.stage in (Session.IDLE, Session.CAT):
            return                                                 # the caller has caught up (or nothing was queued)
        e = self.engine
        rc = e.lib.bf_engine_rollback(e.e, self._stp)
        if rc:
            e._check(rc, "bf_engine_rollback")
        ops.Profile.launches += 1
        e._state_fresh = False
        if self.stage == Session.CORR:
            self._phase(_lib.PH_CORR | _lib.PH_COMPACT)
            if self.valid_called:
                self._phase(_lib.PH_VALID)

    def new_container(self, kind, rows, image_size):
        from .instances import Instances3D
        c = Instances3D.__new__(Instances3D)
        object.__setattr__(c, "_image_size", image_size)
        object.__setattr__(c, "_fields", self.fields(kind, rows))
        return c

    def fields(self, kind, rows) -> EngineFields:
        if kind == "store":
            return EngineFields(self, kind, rows, _STORE_NATIVE + tuple(self.store_chunks))
        extras = self.cat_extras if kind == "cat" else self.map_extras
        f = EngineFields(self, kind, rows, _MAP_NATIVE)
        for k, v in extras.items():
            dict.__setitem__(f, k, v)
        return f

    def view(self, f: EngineFields, k):
        """One field of a handed-out container as a view of the engine state (same layout FusionEngine.export produces).
        Engine-held fields are CUDA views whatever device the caller kept them on (demo.py:216-219 keeps the bookkeeping
        fields on the host); leaving the fast path moves them back."""
        from .boxes import GeneralInstance3DBoxes
        e = self.engine
        if f.kind != "store" or k not in self.store_chunks:
            self.settle()                                          # engine buffers are about to be read
        if f.kind == "store":
            M, st = f.rows, e.store
            if k == "scores":
                return st["scores"][:M, 0]
            if k == "pred_boxes_3d":
                return GeneralInstance3DBoxes._wrap(st["tensor"][:M], st["R"][:M].view(M, 3, 3))
            if k == "cam_pose":
                return st["pose"][:M].view(M, 4, 4)
            if k == "projected_boxes":
                return st["uv"][:M].view(M, 8, 2)
            chunks = self.store_chunks[k]
            if len(chunks) > 1:
                merged = chunks[0]
                for c in chunks[1:]:
                    merged = _cat_values(merged, c)
                chunks[:] = [merged]
            return chunks[0]
        N, mp = f.n_rows(), e.map
        if k == "scores":
            return mp["scores"][:N]
        if k == "pred_boxes":
            return mp["box2d"][:N]
        if k == "pred_proj_xy":
            return mp["projxy"][:N]
        if k == "pred_boxes_3d":
            return GeneralInstance3DBoxes._wrap(mp["tensor"][:N], mp["R"][:N].view(N, 3, 3))
        if k == "cam_pose":
            return mp["pose"][:N].view(N, 4, 4)
        if k == "frame_id":
            return mp["frame_id"][:N].to(torch.int64)
        if k == "init_id":
            return mp["init_id"][:N].to(torch.int64)
        if k == "valid_num":
            return mp["valid"][:N]
        if k == "projected_boxes":
            return mp["uv"][:N].view(N, 8, 2)
        raise KeyError(k)

    # ---- entering: import plain containers into the engine ---------------------------------------------------------
    @staticmethod
    def importable(A, P, bm) -> bool:
        fa, fp = _plain_fields(A), _plain_fields(P)
        if not all(k in fa for k in _MAP_NATIVE) or not all(k in fp for k in _MAP_NATIVE):
            return False
        t = fa["pred_boxes_3d"].tensor
        if not (isinstance(t, torch.Tensor) and t.is_cuda and fp["pred_boxes_3d"].tensor.is_cuda and fa["scores"].is_cuda):
            return False
        return len(bm._fusion_list) == len(A) and len(bm._fusion_flag) == len(P)

    def import_state(self, A, P, bm):
        e = self.engine
        N, M = len(A), len(P)
        if N > e.ncap - e.max_det or M > e.mcap - e.max_det:
            raise RuntimeError("fast path: engine capacity too small for the imported state")
        fa, fp = _plain_fields(A), _plain_fields(P)
        mp, st = e.map, e.store
        with torch.cuda.device(self.dev):
            mp["tensor"][:N].copy_(fa["pred_boxes_3d"].tensor); mp["R"][:N].copy_(fa["pred_boxes_3d"].R.reshape(N, 9))
            mp["scores"][:N].copy_(fa["scores"]); mp["box2d"][:N].copy_(fa["pred_boxes"]); mp["projxy"][:N].copy_(fa["pred_proj_xy"])
            mp["pose"][:N].copy_(fa["cam_pose"].reshape(N, 16)); mp["uv"][:N].copy_(fa["projected_boxes"].reshape(N, 16))
            mp["valid"][:N].copy_(fa["valid_num"]); mp["init_id"][:N].copy_(fa["init_id"]); mp["frame_id"][:N].copy_(fa["frame_id"])
            st["tensor"][:M].copy_(fp["pred_boxes_3d"].tensor); st["R"][:M].copy_(fp["pred_boxes_3d"].R.reshape(M, 9))
            st["scores"][:M, 0].copy_(fp["scores"]); st["uv"][:M].copy_(fp["projected_boxes"].reshape(M, 16))
            st["pose"][:M].copy_(fp["cam_pose"].reshape(M, 16))
            fl, ln, _ = bm.pack_lists(N)
            mp["fl"][:N].copy_(torch.from_numpy(fl)); mp["flen"][:N].copy_(torch.from_numpy(ln))
            e.fflag[:M].copy_(torch.from_numpy(np.asarray(bm._fusion_flag, dtype=np.int32)))
            af = bm._already_fusion
            F = len(af)
            if F > e.fcap:
                raise RuntimeError("fast path: fused_capacity too small")
            if F:
                lens = np.fromiter(map(len, af), dtype=np.int32, count=F)
                if lens.max() > ops.FUSION_CAP:
                    raise RuntimeError("fast path: an already_fusion entry exceeds the device list capacity")
                tab = np.zeros((F, ops.FUSION_CAP), dtype=np.int32)
                for i, l in enumerate(af):
                    tab[i, :len(l)] = l
                hs = np.array([list_hash(l) for l in af], dtype=np.uint64).view(np.int64)
                e.fused["lists"][:F].copy_(torch.from_numpy(tab)); e.fused["len"][:F].copy_(torch.from_numpy(lens))
                e.fused["hash"][:F].copy_(torch.from_numpy(hs))
            e.fused["count"].fill_(F)
            if e.stream is not None:
                e.stream.wait_stream(torch.cuda.current_stream(self.dev))
        self._stp = e._st()
        e._check(e.lib.bf_engine_set_counts(e.e, N, M, self._stp), "bf_engine_set_counts")
        e._state_fresh = False
        e.M, e._n_ub = M, N
        self.N, self.M = N, M
        self.map_extras = {k: v for k, v in fa.items() if k not in _MAP_NATIVE}
        self.store_chunks = {k: [v] for k, v in fp.items() if k not in _STORE_NATIVE}
        self.map_dev = {k: fa[k].device for k in _MAP_NATIVE if isinstance(fa[k], torch.Tensor)}
        self.store_dev = {k: fp[k].device for k in _STORE_NATIVE if isinstance(fp[k], torch.Tensor)}
        self.image_size = A.image_size
        # the caller's own containers become views of the engine state (in-place edits stay coherent)
        for c, kind, rows in ((A, "map", N), (P, "store", M)):
            object.__setattr__(c, "_fields", self.fields(kind, rows))
        self.map_c, self.store_c, self.cat_c = A, P, None
        self.stage = Session.IDLE
        self.ahead, self.valid_called = False, False
        self.lists_host, self.lists_given = True, None

    # ---- leaving: export the engine state into plain containers ----------------------------------------------------
    def pull_lists(self, bm):
        """BoxManager's Python lists from the device (the device is ahead while a session runs)."""
        self.settle()
        e = self.engine
        if self.stage in (Session.CAT, Session.NMS):
            rows = self.N + self.n
        else:
            rows = self.N if self.N is not None else e.N
        s = e.state()                                              # synchronises the engine's stream
        M = self.M + (self.n if self.stage != Session.IDLE else 0)
        mp = e.map
        flen = mp["flen"][:rows].cpu().numpy()
        fl = mp["fl"][:rows].cpu().numpy()
        bm._fusion_list = [[int(x) for x in fl[i, :flen[i]]] for i in range(rows)]
        bm._fusion_flag = [int(x) for x in e.fflag[:M].cpu().numpy()]
        F = int(e.fused["count"].item())
        al, aln = e.fused["lists"][:F].cpu().numpy(), e.fused["len"][:F].cpu().numpy()
        bm._already_fusion = [[int(x) for x in al[i, :aln[i]]] for i in range(F)]
        bm._fused_set, bm._fused_n = set(), -1
        bm.last_fusion_frame = [[0] for _ in range(M)]
        ops.Profile.d2h_bytes += 4 * (rows * (ops.FUSION_CAP + 1) + M + F * (ops.FUSION_CAP + 1))
        del s

    def lists_for_caller(self, bm):
        """The caller looks at fusion_list / fusion_flag / already_fusion while the session runs."""
        if not self.lists_host:
            self.pull_lists(bm)
            self.lists_host = True
            self.lists_given = ([list(l) for l in bm._fusion_list], list(bm._fusion_flag), [list(l) for l in bm._already_fusion])

    def lists_untouched(self, bm) -> bool:
        g = self.lists_given
        return g is None or (g[0] == bm._fusion_list and g[1] == bm._fusion_flag and g[2] == bm._already_fusion)

    def detach(self, bm):
        """Leave the fast path: everything the session holds becomes plain state again."""
        if bm is None or bm._session is not self:
            return
        e = self.engine
        self._stp = e._st()
        self.settle()
        edited = not self.lists_untouched(bm)
        if self.stage == Session.CORR:                             # close the keyframe on the device too (row counters)
            self._phase(_lib.PH_FINISH)
            self.M += self.n
            self.stage = Session.IDLE
        if not edited and not self.lists_host:
            self.pull_lists(bm)
        if self.stage == Session.IDLE and self.pending_new is not None:   # rows init_new_predictions announced, not yet ingested
            n_new, m0 = self.pending_new
            for i in range(n_new):
                bm._fusion_list.append([i + m0]); bm.last_fusion_frame.append([0]); bm._fusion_flag.append(0)
        self.pending_new = None
        e.state()                                                  # everything issued so far has completed
        for c in (self.map_c, self.store_c, self.cat_c):
            if c is not None and isinstance(c._fields, EngineFields):
                f = c._fields
                f.fill()
                devs = self.store_dev if f.kind == "store" else self.map_dev
                plain = {}
                for k, v in dict.items(f):                         # independent copies, on the devices the caller kept the fields on
                    d = devs.get(k)
                    if isinstance(v, torch.Tensor):
                        v = v.to(d) if (d is not None and d != v.device) else v.clone()
                    elif hasattr(v, "clone"):
                        v = v.clone()
                    plain[k] = v
                object.__setattr__(c, "_fields", plain)
        bm._session = None
        self.stage = Session.IDLE
        self.map_c = self.store_c = self.cat_c = self.pred = None

    # ---- the calls ---------------------------------------------------------------------------------------------------
    def try_cat(self, lst):
        if len(lst) != 2:
            return None
        A, B = lst
        bm = self.bm()
        if bm is None or bm._session is not self:
            return None
        if A is self.store_c:                                      # cat([per_frame_ins, pred_instances]) (demo.py:254)
            if B is not self.pred or self.stage != Session.CAT:
                return None
            f = _plain_fields(B)
            for k, chunks in self.store_chunks.items():
                chunks.append(f[k])
            self.store_c = self.new_container("store", self.M + self.n, A.image_size)
            return self.store_c
        if A is not self.map_c:
            return None
        if self.stage == Session.CORR:                             # boxfusion() was not called for the last keyframe (cfg / no new box)
            if self.ahead and ((self.use_fusion and self.any_new) or (self.check_valid and self.any_new and not self.valid_called)):
                self.settle()                                      # the engine fused / dropped what the caller never asked for: undo
            if self.ahead:
                self.ahead = False                                 # what ran ahead is exactly what the caller's calls amount to
            else:
                self._phase(_lib.PH_FINISH)
            self.engine.M += self.n
            self.engine.count += 1
            self.M += self.n
            self.stage, self.n, self.pred = Session.IDLE, 0, None
        if self.stage != Session.IDLE or not self.lists_untouched(bm) or not isinstance(A._fields, EngineFields):
            return None
        fb = B._fields
        if isinstance(fb, EngineFields) or not all(k in fb for k in _MAP_NATIVE):
            return None
        proj = getattr(B, "_bf_proj", None)
        boxes = fb["pred_boxes_3d"]
        n = len(boxes)
        e = self.engine
        if proj is None or n < 1 or n > e.max_det or not boxes.tensor.is_cuda or boxes.tensor.device != self.dev:
            return None
        t, R, sc, b2, pxy, uv = boxes.tensor, boxes.R, fb["scores"], fb["pred_boxes"], fb["pred_proj_xy"], fb["projected_boxes"]
        for x in (t, R, sc, b2, pxy, uv):
            if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
                return None
        if self.N is None:
            self.N = e.N
        if self.N + n > e.ncap or self.M + n > e.mcap:
            return None
        extras = [k for k in fb if k not in _MAP_NATIVE]
        if set(self.map_extras) - set(extras):
            return None
        K, H, W, pose = proj
        fid = fb["frame_id"]
        frame_id = int(fid[0]) if isinstance(fid, torch.Tensor) else int(np.asarray(fid).reshape(-1)[0])
        # header of the keyframe: n, frame id, intrinsics, pose, np.linalg.inv(pose) for the correspondence projection
        # (instances.py:680); the observation projection was done by the caller's project_3d_boxes
        pose = np.ascontiguousarray(pose.detach().cpu().numpy() if isinstance(pose, torch.Tensor) else pose, dtype=np.float32).reshape(4, 4)
        hdr = np.zeros(KF_HEADER, dtype=np.float32)
        hi = hdr.view(np.int32)
        hi[0], hi[1] = n, frame_id
        K3 = np.asarray(K, dtype=np.float32)
        hdr[2:8] = (K3[0, 0], K3[1, 1], K3[0, 2], K3[1, 2], float(W), float(H))
        hdr[8:24] = pose.reshape(-1)
        hdr[40:56] = np.linalg.inv(pose).astype(np.float32).reshape(-1)
        st = self._stp = e._st()
        if e.stream is not None:
            e.stream.wait_stream(torch.cuda.current_stream(self.dev))
        rc = e.lib.bf_engine_ingest_world(e.e, hdr.ctypes.data, t.data_ptr(), R.data_ptr(), sc.data_ptr(), b2.data_ptr(),
                                          pxy.data_ptr(), uv.data_ptr(), n, st)
        if rc:
            e._check(rc, "bf_engine_ingest_world")
        ops.Profile.launches += 1
        ops.Profile.h2d_bytes += 4 * KF_HEADER
        ops.Profile.calls["bf_engine_ingest_world"] = ops.Profile.calls.get("bf_engine_ingest_world", 0) + 1
        e._state_fresh = False
        self.hdr, self.n, self.pred, self.frame_id = hdr, n, B, frame_id
        self.pending_new = None
        self.valid_called = False
        self.cat_extras = {k: _cat_values(v, fb[k]) for k, v in self.map_extras.items()}
        self.stage = Session.CAT
        self.lists_host, self.lists_given = False, None
        self.cat_c = self.new_container("cat", self.N + n, A.image_size)
        return self.cat_c

    def try_nms(self, A, threshold, bm):
        if A is not self.cat_c or self.stage != Session.CAT or bm._session is not self:
            return None
        from . import instances as inst_mod
        if float(threshold) != self.key[0] or inst_mod.IOU_MODE != self.iou_mode or cfg_key(self.cfg) != self.key:
            return None
        rows = self.N + self.n
        if RUN_AHEAD:
            self._run_ahead(rows)                                  # NMS, flags -> pinned, snapshot, correspondence, flags, rest of the keyframe
            k, sflags, _ = self._wait_flags(0, rows, True)
            keep, succ = np.nonzero(k)[0], np.nonzero(sflags)[0]
        else:
            self._phase(_lib.PH_NMS)
            self._read_flags(rows, True)
            keep = np.nonzero(self._keep[:rows])[0]
            succ = np.nonzero(self._succ[:rows])[0]
        self.last_keep = keep
        self.stage = Session.NMS
        return keep.tolist(), succ.tolist()

    def _after_assoc(self, image_size):
        """PH_CORR, the keep flags, PH_COMPACT -> (container for all_pred_box[keep_idx], keep_idx)."""
        rows = self.N + self.n
        if self.ahead:
            k, _, st = self._wait_flags(1, rows, False)            # queued behind spatial_association; usually long done
            keep_idx = np.nonzero(k)[0]
            self.any_new = int(st.any_new)
        else:
            self._phase(_lib.PH_CORR)
            self._read_flags(rows, False)
            keep_idx = np.nonzero(self._keep[:rows])[0]
            self.any_new = int(self.engine._state.any_new)
            self._phase(_lib.PH_COMPACT)
        self.map_extras = {k: _index_values(v, keep_idx) for k, v in self.cat_extras.items()}
        self.cat_extras = {}
        self.N = int(len(keep_idx))
        self.stage = Session.CORR
        self.map_c = self.new_container("map", self.N, image_size)
        self.cat_c = None
        return self.map_c, keep_idx

    def try_corr(self, cfg, bm, pred_instances, all_pred_box, all_poses, frame_id, mask, intrinsic, threshold, H, W):
        if all_pred_box is not self.cat_c or self.stage != Session.NMS or bm._session is not self or pred_instances is not self.pred:
            return None
        K, Hh, Wh, _ = self.pred._bf_proj
        Ki = intrinsic.detach().cpu().numpy() if isinstance(intrinsic, torch.Tensor) else np.asarray(intrinsic)
        if (float(threshold) != self.key[1] or float(H) != float(Hh) or float(W) != float(Wh)
                or not np.array_equal(np.asarray(Ki, dtype=np.float32)[:3, :3], np.asarray(K, dtype=np.float32)[:3, :3])
                or cfg_key(cfg) != self.key
                or not np.array_equal(np.asarray(mask), self.last_keep)):
            return None
        c, keep_idx = self._after_assoc(all_pred_box.image_size)
        return c, all_poses[keep_idx], keep_idx

    def try_getitem(self, A, item):
        """`all_pred_box[mask]` when no new box survived nms_3d (demo.py:325)."""
        if A is not self.cat_c or self.stage != Session.NMS:
            return None
        bm = self.bm()
        if bm is None or bm._session is not self:
            return None
        try:
            same = np.array_equal(np.asarray(item), self.last_keep)
        except Exception:
            same = False
        if not same:
            return None
        return self._after_assoc(A.image_size)[0]

    def try_check_valid(self, bm, all_pred_box, count, gap):
        if (all_pred_box is not self.map_c or self.stage != Session.CORR or bm._session is not self or self.map_extras
                or not self.check_valid or int(gap) != self.gap or int(count) != self.frame_id):
            return None
        if not self.ahead:
            self._phase(_lib.PH_VALID)
        self.valid_called = True
        self.N = None                                             # known on the device only
        self.map_c = self.new_container("map", None, all_pred_box.image_size)
        return self.map_c

    def try_boxfusion(self, fuser, all_pred_box, per_frame_box, bm, beta):
        if (all_pred_box is not self.map_c or per_frame_box is not self.store_c or self.stage != Session.CORR or bm._session is not self
                or beta != 0.9 or not fuser.early_stop or not self.use_fusion or cfg_key(fuser.cfg) != self.key):
            return None
        K, H, W, _ = self.pred._bf_proj
        if (float(fuser.H) != float(H) or float(fuser.W) != float(W)
                or not np.array_equal(np.asarray(fuser.K, dtype=np.float32)[:3, :3], np.asarray(K, dtype=np.float32)[:3, :3])):
            return None
        if self.ahead and self.check_valid and self.any_new and not self.valid_called:
            return None                                            # the engine dropped stale rows the caller never asked to drop
        if self.ahead:
            self.ahead = False                                     # the engine did all of this behind spatial_association
        else:
            self._phase(_lib.PH_FUSE | _lib.PH_FINISH)
        e = self.engine
        e.M += self.n
        e._n_ub = (self.N if self.N is not None else e._n_ub)
        e.count += 1
        self.M += self.n
        self.stage, self.n, self.pred = Session.IDLE, 0, None
        fuser.last_iters = None
        return True


def session_of(bm) -> Optional[Session]:
    return getattr(bm, "_session", None) if ENABLED else None


def leave(bm):
    s = getattr(bm, "_session", None)
    if s is not None:
        s.detach(bm)


def find_session(*containers) -> Optional[Session]:
    if not ENABLED:
        return None
    for c in containers:
        f = getattr(c, "_fields", None)
        if isinstance(f, EngineFields):
            return f.sess
    return None
