"""Multi-GPU checks (need >= 2 visible GPUs; run with `gpurun --gpus 2`): NCCL broadcast of the rolling
terrain prior on the handles' own device memory, then independent clouds per rank against it."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    from groundgrid_b200 import capi, prior, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dim, res = 99.0, 0.33
    scene = synth.make_scene(seed=9)
    g = capi.GroundGridB200(dim, res, device=rank, n_slots=1, max_points=131072, full_layers=False)
    g.init_map(0.0, 0.0, 0.0)
    if rank == 0:
        for k in range(3):
            pts, org = synth.scan_64(scene, ego_xy=(0.7 * k, 0.0), seed=k)
            if k:
                g.update_pose(0.7 * k, 0.0, synth.base_from_map(0.7 * k, 0.0))
            g.filter_cloud(pts, org, 0.0)
    xy = prior.broadcast_prior(g, src=0)
    pts, org = synth.scan_64(scene, ego_xy=(1.4, 0.0), seed=100 + rank)
    labels = g.filter_cloud(pts, org, 0.0)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), labels=labels, ground=g.layer("ground"), pos=xy)
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_prior_broadcast_two_gpus(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from groundgrid_b200 import synth
    from oracle import Oracle

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    dim, res = 99.0, 0.33
    scene = synth.make_scene(seed=9)
    o = Oracle(dim, res)
    o.init_map(0.0, 0.0, 0.0)
    for k in range(3):
        pts, org = synth.scan_64(scene, ego_xy=(0.7 * k, 0.0), seed=k)
        if k:
            o.update(0.7 * k, 0.0, synth.base_from_map(0.7 * k, 0.0))
        o.filter_cloud(pts, org, 0.0, threads=1)
    G, C = o.layer("ground"), o.layer("groundpatch")
    for rank in range(2):
        r = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(r["pos"], o.position())
        o.set_layer("ground", G)
        o.set_layer("groundpatch", C)
        pts, org = synth.scan_64(scene, ego_xy=(1.4, 0.0), seed=100 + rank)
        want, _, _ = o.filter_cloud(pts, org, 0.0, threads=1)
        assert np.array_equal(r["labels"], want), f"rank {rank}"
        assert np.array_equal(r["ground"], o.layer("ground")), f"rank {rank}"
