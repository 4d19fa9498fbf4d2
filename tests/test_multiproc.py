"""CPU test of the N>1 host logic (world_size 2, gloo): sequence sharding + the final map all_gather.
The per-rank fusion itself runs the CPU oracle port here (the CUDA product needs a GPU); what is under test is
boxfusion_b200/sharding.py, the code bench.py uses under torchrun."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_SEQ, N_FRAMES = 4, 3


def _run_sequence(seed):
    from boxfusion_b200.driver import FusionSession
    from boxfusion_b200.synthetic import SyntheticScene, make_cfg, make_pst
    from boxfusion_b200.sharding import map_rows
    from oracle import port
    port.IOU_BACKEND = "c"
    scene = SyntheticScene(n_objects=12, seed=seed, max_det=6)
    cfg = make_cfg("ca1m", pst_path=make_pst(64, seed=0), pst_size=64)
    cfg["box_fusion"]["iters"] = 3
    sess = FusionSession(port, cfg)
    for k in range(N_FRAMES):
        sess.step(scene.keyframe(k))
    return map_rows(sess.all_pred_box)


def _worker(rank, world, port_no, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from boxfusion_b200.sharding import gather_maps, shard_sequences
    mine = shard_sequences(N_SEQ, rank, world)
    rows = torch.cat([_run_sequence(100 + s) for s in mine], dim=0)
    maps = gather_maps(rows)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), torch.cat(maps, dim=0).numpy())
        np.save(os.path.join(out_dir, "sizes.npy"), np.array([m.shape[0] for m in maps]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather(tmp_path):
    from boxfusion_b200.sharding import shard_sequences
    assert shard_sequences(5, 0, 2) == [0, 2, 4] and shard_sequences(5, 1, 2) == [1, 3]
    assert sorted(shard_sequences(64, r, 8)[0] for r in range(8)) == list(range(8))
    port_no = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port_no, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npy")
    sizes = np.load(tmp_path / "sizes.npy")
    # single-process reference: rank 0 owns sequences 0,2; rank 1 owns 1,3
    exp = torch.cat([_run_sequence(100 + s) for s in (0, 2, 1, 3)], dim=0).numpy()
    assert got.shape == exp.shape and np.array_equal(got, exp)
    assert sizes.sum() == exp.shape[0] and len(sizes) == 2


# ---- sharding inside one step: box-sharded refinement and row-block IoU matrix (SURVEY 8(e) axes 2, 3) ------------------
def _refine_inputs():
    from boxfusion_b200.synthetic import make_cfg, make_pst, refine_problem
    from oracle import port, refine_oracle as ro
    B, V, P = 5, 3, 64
    prob = refine_problem(B, V, seed=4)
    W, H = prob["size"]
    cfg = make_cfg("ca1m", pst_path=None, pst_size=P)
    cfg["box_fusion"]["iters"] = 4
    uv = []
    for b in range(B):
        ins = port.Instances3D((H, W))
        ins.pred_boxes_3d = port.GeneralInstance3DBoxes(torch.from_numpy(prob["tensor"][b]), torch.from_numpy(prob["R"][b]))
        ins.cam_pose = torch.from_numpy(prob["poses"][b])
        ins.project_3d_boxes(prob["K"], H=H, W=W)
        uv.append(ins.projected_boxes.numpy().reshape(V, 16))
    cs = ro.make_cfg_struct(cfg, H, W)
    return prob, np.stack(uv), make_pst(P, seed=1), ro.K16_from_K3(prob["K"]), cs, B, V


def _oracle_refine_fn(K16, cs):
    """ops.refine's contract (CSR in, (out, updated, iters) out) served by the CPU oracle: what is under test is the
    partitioning and the gathers of sharding.refine_sharded, not the arithmetic."""
    from oracle import refine_oracle as ro

    def fn(pst, t, R, s, uv, po, off, idx, rcfg):
        out, upd, its = [], [], []
        for b in range(len(off) - 1):
            v = np.asarray(idx[off[b]:off[b + 1]])
            u, o6, n_it, _ = ro.refine_box(t[v], R[v], s[v], uv[v], po[v], pst, K16, cs)
            out.append(o6 if u else np.zeros(6, np.float32)); upd.append(int(u)); its.append(n_it)
        return (torch.from_numpy(np.asarray(out, np.float32).reshape(-1, 6)), torch.tensor(upd, dtype=torch.int32),
                torch.tensor(its, dtype=torch.int32))
    return fn


def _worker_step(rank, world, port_no, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from boxfusion_b200.sharding import iou3d_matrix_sharded, refine_sharded
    from oracle import port
    prob, uv, pst, K16, cs, B, V = _refine_inputs()
    off = np.arange(B + 1, dtype=np.int32) * V
    idx = np.arange(B * V, dtype=np.int32)
    out, upd, its = refine_sharded(pst, prob["tensor"].reshape(-1, 6), prob["R"].reshape(-1, 3, 3), prob["scores"].reshape(-1),
                                   uv.reshape(-1, 16), prob["poses"].reshape(-1, 4, 4), off, idx, None,
                                   refine_fn=_oracle_refine_fn(K16, cs))
    ca = port.GeneralInstance3DBoxes(torch.from_numpy(prob["tensor"][:, 0]), torch.from_numpy(prob["R"][:, 0])).corners
    cb = port.GeneralInstance3DBoxes(torch.from_numpy(prob["tensor"][:, 1]), torch.from_numpy(prob["R"][:, 1])).corners
    port.IOU_BACKEND = "c"
    iou = iou3d_matrix_sharded(ca, cb, 0, iou_fn=lambda a, b, mode: torch.from_numpy(
        np.stack([port.calculate_obb_iou(x.numpy(), b.numpy()) for x in a]).reshape(a.shape[0], b.shape[0])))
    if rank == 1:                     # every rank holds the full result: check the non-zero rank's copy
        np.savez(os.path.join(out_dir, "step.npz"), out=out.numpy(), upd=upd.numpy(), its=its.numpy(), iou=iou.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_box_sharded_refine_and_row_sharded_iou(tmp_path):
    from boxfusion_b200.sharding import block_range
    assert [block_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [block_range(3, r, 8) for r in range(8)] == [(0, 1), (1, 2), (2, 3)] + [(3, 3)] * 5
    covered = [i for r in range(8) for i in range(*block_range(128, r, 8))]
    assert covered == list(range(128))
    port_no = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_step, args=(2, port_no, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "step.npz")
    from oracle import port
    prob, uv, pst, K16, cs, B, V = _refine_inputs()
    fn = _oracle_refine_fn(K16, cs)
    off = np.arange(B + 1, dtype=np.int32) * V
    out, upd, its = fn(pst, prob["tensor"].reshape(-1, 6), prob["R"].reshape(-1, 3, 3), prob["scores"].reshape(-1),
                       uv.reshape(-1, 16), prob["poses"].reshape(-1, 4, 4), off, np.arange(B * V, dtype=np.int32), None)
    assert np.array_equal(got["out"].view(np.uint32), out.numpy().view(np.uint32))
    assert np.array_equal(got["upd"], upd.numpy()) and np.array_equal(got["its"], its.numpy())
    port.IOU_BACKEND = "c"
    ca = port.GeneralInstance3DBoxes(torch.from_numpy(prob["tensor"][:, 0]), torch.from_numpy(prob["R"][:, 0])).corners.numpy()
    cb = port.GeneralInstance3DBoxes(torch.from_numpy(prob["tensor"][:, 1]), torch.from_numpy(prob["R"][:, 1])).corners.numpy()
    exp = np.stack([port.calculate_obb_iou(x, cb) for x in ca])
    assert got["iou"].shape == (B, B) and np.array_equal(got["iou"], exp)


# ---- row-sharded NMS: edge lists gathered, greedy scan replicated (SURVEY 8(e) axis 3) ------------------------------------
def _nms_inputs():
    from boxfusion_b200.synthetic import map_and_detections
    from oracle import port
    (mt, mR, ms), (dt, dR, ds) = map_and_detections(40, 14, seed=6, tilt_noise=0.01)
    t, R, s = np.concatenate([mt, dt]), np.concatenate([mR, dR]), np.concatenate([ms, ds])
    corners = port.GeneralInstance3DBoxes(torch.from_numpy(t), torch.from_numpy(R)).corners.numpy()
    order = np.argsort(-s, kind="stable")
    rank = np.empty_like(order); rank[order] = np.arange(len(order))
    return corners, s, rank


def _edges_of_rows(corners, rank, lo, hi, thr=0.1, cap=64):
    """the contract of ops.nms3d_edges served by the CPU oracle: keys rank_lo << 32 | rank_hi, empty slots -1"""
    from oracle import port
    port.IOU_BACKEND = "c"
    keys = []
    for a in range(lo, hi):
        if a + 1 < len(corners):
            iou = port.calculate_obb_iou(corners[a], corners[a + 1:])
            for j in np.nonzero(iou > thr)[0]:
                b = a + 1 + int(j)
                r0, r1 = sorted((int(rank[a]), int(rank[b])))
                keys.append((r0 << 32) | r1)
    out = np.full(cap, -1, dtype=np.int64)
    out[: len(keys)] = keys
    return torch.from_numpy(out), torch.zeros(1, dtype=torch.int32)


def _worker_nms(rank, world, port_no, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from boxfusion_b200.sharding import nms3d_sharded
    corners, s, rk = _nms_inputs()
    got = {}

    def greedy(edges):                                  # what reaches the replicated greedy scan is the thing under test
        got["edges"] = edges.numpy().copy()
        n = len(corners)
        return torch.zeros(n, dtype=torch.int32), torch.zeros(n, dtype=torch.int32), torch.zeros(1, dtype=torch.int32)
    nms3d_sharded(torch.from_numpy(corners), None, None, None, None, None, None, None, 0.1, 0.8, 30.0,
                  edges_fn=lambda lo, hi: _edges_of_rows(corners, rk, lo, hi), greedy_fn=greedy)
    np.save(os.path.join(out_dir, f"edges{rank}.npy"), got["edges"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_sharded_nms_edges(tmp_path):
    from boxfusion_b200.sharding import pair_row_ranges
    for n, world in ((54, 2), (4352, 8), (5, 8), (1, 2), (0, 2)):
        rr = pair_row_ranges(n, world)
        assert rr[0][0] == 0 and rr[-1][1] == n and all(a[1] == b[0] for a, b in zip(rr, rr[1:]))
        pairs = [sum(n - 1 - a for a in range(lo, hi)) for lo, hi in rr]
        assert sum(pairs) == n * (n - 1) // 2
        if n == 4352:
            assert max(pairs) - min(pairs) <= 2 * n              # balanced by pair count (to within two rows), not by row count
    port_no = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker_nms, args=(2, port_no, str(tmp_path)), nprocs=2, join=True)
    corners, s, rk = _nms_inputs()
    full, _ = _edges_of_rows(corners, rk, 0, len(corners), cap=128)
    want = np.sort(full.numpy()[full.numpy() >= 0])
    for r in range(2):                                          # every rank feeds the same edge set to its greedy scan
        e = np.load(tmp_path / f"edges{r}.npy")
        assert e.shape == (128,) and np.array_equal(np.sort(e[e >= 0]), want) and len(want) > 5
