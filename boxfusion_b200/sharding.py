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
