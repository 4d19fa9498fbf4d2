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
