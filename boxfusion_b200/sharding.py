"""Multi-GPU plumbing of the fusion path: independent sequences are sharded round-robin over ranks
(one process per GPU, torch.distributed), no data-path collective; the only exchange is a final
all_gather of the per-rank maps.  Works with NCCL (CUDA tensors) and gloo (CPU)."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_sequences(n_sequences: int, rank: int, world: int) -> List[int]:
    """Indices of the sequences rank `rank` owns (round-robin)."""
    return list(range(rank, n_sequences, world))


def map_rows(all_pred_box) -> torch.Tensor:
    """[N,15] float32 rows (x,y,z,l,h,w,R[9]) of a map held in an Instances3D."""
    b = all_pred_box.pred_boxes_3d
    return torch.cat([b.tensor, b.R.reshape(-1, 9)], dim=1).contiguous()


def gather_maps(rows: torch.Tensor, group=None) -> List[torch.Tensor]:
    """all_gather of variable-length [N_i,15] maps: sizes first, then zero-padded rows; returns the list per rank."""
    world = dist.get_world_size(group)
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    nmax = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((nmax, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    pad[: rows.shape[0]] = rows
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[: int(s.item())] for o, s in zip(out, sizes)]


# ---- sharding inside one step (SURVEY.md section 8(e) axes 2 and 3) ----------------------------------------------------
# Map boxes of one refinement call are independent, and so are the rows of an IoU matrix: rank r takes a contiguous block,
# computes it with the single-device library, and one all_gather of the (small) results closes the step.  No collective
# inside the kernels - there is no exchange step in the algorithm.

def block_range(n: int, rank: int, world: int):
    """Contiguous block [lo, hi) of n items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_blocks(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all_gather of the ranks' contiguous blocks (block_range order) -> the full [n_total, ...] tensor on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nmax = max(block_range(n_total, r, world)[1] - block_range(n_total, r, world)[0] for r in range(world))
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    sizes = [block_range(n_total, r, world)[1] - block_range(n_total, r, world)[0] for r in range(world)]
    assert sizes[rank] == local.shape[0]
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def refine_sharded(pst, per_xyzlhw, per_R, per_scores, per_uv, per_poses, view_offsets, view_index, rcfg,
                   refine_fn=None, group=None):
    """BoxFusion.boxfusion's optimiser (box_fusion.py:651-721) for B boxes, box-sharded over the ranks: every rank holds the
    whole observation store, refines its block of the CSR with `refine_fn` (default: ops.refine, one bf_refine launch) and
    all_gathers (out_xyzlhw [B,6], updated [B], iters [B])."""
    import numpy as np
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    off_h = view_offsets.cpu().numpy() if isinstance(view_offsets, torch.Tensor) else np.asarray(view_offsets)
    B = len(off_h) - 1
    lo, hi = block_range(B, rank, world)
    off = (off_h[lo:hi + 1] - off_h[lo]).astype(np.int32)
    idx = view_index[int(off_h[lo]):int(off_h[hi])]
    if hi > lo:
        if refine_fn is None:
            from . import ops
            out, upd, its, _, status = ops.refine(pst, per_xyzlhw, per_R, per_scores, per_uv, per_poses, off, idx, rcfg,
                                                  max_views=int(np.max(np.diff(off))))
            if int(status.item()) != 0:
                raise RuntimeError("bf_refine: capacity exceeded (views per box or polygon candidates)")
        else:
            out, upd, its = refine_fn(pst, per_xyzlhw, per_R, per_scores, per_uv, per_poses, off, idx, rcfg)
    else:
        dev = pst.device if isinstance(pst, torch.Tensor) else "cpu"
        out, upd, its = (torch.zeros((0, 6), dtype=torch.float32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev),
                         torch.zeros(0, dtype=torch.int32, device=dev))
    return gather_blocks(out, B, group), gather_blocks(upd, B, group), gather_blocks(its, B, group)


def iou3d_matrix_sharded(cornersA: torch.Tensor, cornersB: torch.Tensor, mode: int = 0, iou_fn=None, group=None) -> torch.Tensor:
    """calculate_obb_iou over all pairs (instances.py:106-125), row-block sharded: rank r computes its rows of the [M,N]
    float64 matrix with `iou_fn` (default: ops.iou3d_matrix) and the blocks are all_gathered."""
    if iou_fn is None:
        from . import ops
        iou_fn = ops.iou3d_matrix
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    M = cornersA.shape[0]
    lo, hi = block_range(M, rank, world)
    if hi > lo:
        local = iou_fn(cornersA[lo:hi], cornersB, mode)
    else:
        local = torch.zeros((0, cornersB.shape[0]), dtype=torch.float64, device=cornersB.device)
    return gather_blocks(local, M, group)


def pair_row_ranges(n: int, world: int):
    """Row blocks [lo, hi) of the pair triangle {(a, b): a < b < n} with (nearly) equal pair counts: row a has n-1-a pairs."""
    total = n * (n - 1) // 2
    bounds, a, acc = [0], 0, 0
    for r in range(1, world):
        target = total * r // world
        while a < n and acc + (n - 1 - a) <= target:
            acc += n - 1 - a
            a += 1
        bounds.append(a)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def nms3d_sharded(corners, centers, order, init_id, poses, fusion_list, fusion_len, fusion_flag, iou_threshold, translation_gap,
                  rotation_gap, center_gap=0.5, mode=0, edges_fn=None, greedy_fn=None, group=None):
    """nms_3d + BoxManager.record (instances.py:22-101, box_manager.py:40-88) with the IoU work sharded by rows of the pair
    triangle (SURVEY.md section 8(e) axis 3): every rank finds the over-threshold pairs of its rows (bf_nms3d_edges), ONE
    all_gather of the edge lists - 8 bytes per over-threshold pair, 64 KB at most - and every rank runs the greedy scan with
    record() over the same concatenated list (bf_nms3d_greedy), so the fusion lists end up identical everywhere.
    Returns (keep, success, status) like ops.nms3d.  The dense [N, N] IoU matrix is never formed or moved."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = corners.shape[0]
    lo, hi = pair_row_ranges(n, world)[rank]
    if edges_fn is None or greedy_fn is None:
        from . import ops
        cap = ops.EDGE_CAP // world
        edges_fn = edges_fn or (lambda a, b: ops.nms3d_edges(corners, order, a, b, iou_threshold, mode, edge_cap=cap))
        greedy_fn = greedy_fn or (lambda e: ops.nms3d_greedy(e, centers, order, init_id, poses, fusion_list, fusion_len, fusion_flag,
                                                             translation_gap, rotation_gap, center_gap))
    local, st_local = edges_fn(lo, hi)
    # one all_gather carries the edge slots and, in one extra slot, the rank's status word (BF_ERR_* are negative)
    st_word = st_local.to(torch.int64).reshape(1) if st_local is not None else torch.zeros(1, dtype=torch.int64, device=local.device)
    buf = torch.cat([local.reshape(-1), st_word.to(local.device)])
    out = torch.empty(world * buf.numel(), dtype=buf.dtype, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view(world, -1)
    keep, success, status = greedy_fn(out[:, :-1].reshape(-1).contiguous())
    status = torch.minimum(status, out[:, -1].min().to(torch.int32).reshape(1).to(status.device))   # any rank's overflow reaches everybody
    return keep, success, status
