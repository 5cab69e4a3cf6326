"""world_size-2 checks of the N>1 host logic on the CPU (gloo): stream sharding and the
shared-prior broadcast protocol (the oracle's maps stand in for the device maps)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from groundgrid_b200 import prior, synth
from oracle import Oracle


def test_shard_partitions_everything_once():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            parts = [prior.shard(n, r, world) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dim, res = 33.0, 0.33
    scene = synth.make_scene(seed=5, n_boxes=8, rmin=4.0, rmax=14.0)
    o = Oracle(dim, res)
    o.init_map(0.0, 0.0, 0.0)
    if rank == 0:  # only the owner has seen the earlier scans of the ego frame
        for k in range(3):
            pts, org = synth.lidar_scan(scene, beams=24, az_steps=256, seed=k)
            if k:
                o.update(0.4 * k, 0.0, synth.base_from_map(0.4 * k, 0.0))
            o.filter_cloud(pts, org, 0.0, threads=1)
    # shared-prior step: broadcast ground||groundpatch and the map position from the owner
    n2 = o.n * o.n
    flat = torch.from_numpy(np.concatenate([o.layer("ground").reshape(-1, order="F"), o.layer("groundpatch").reshape(-1, order="F")]))
    pos = torch.from_numpy(o.position().copy())
    prior.broadcast_prior_tensors(flat, pos, src=0)
    if rank != 0:
        # a replica adopts the owner's position by moving its (still pristine) map there, then the layers
        o.update(float(pos[0]), float(pos[1]), synth.base_from_map(float(pos[0]), float(pos[1])))
        assert np.array_equal(o.position(), pos.numpy())
        o.set_layer("ground", flat[:n2].numpy().reshape(o.n, o.n, order="F"))
        o.set_layer("groundpatch", flat[n2:].numpy().reshape(o.n, o.n, order="F"))
    # every rank now evaluates ITS OWN cloud of the shared ego frame against the same prior
    mine = prior.shard(4, rank, world)
    labels = {}
    for c in mine:
        pts, org = synth.lidar_scan(scene, ego_xy=(0.8, 0.0), beams=24, az_steps=256, seed=100 + c)
        G, C = o.layer("ground"), o.layer("groundpatch")
        labels[c], _, _ = o.filter_cloud(pts, org, 0.0, threads=1)
        o.set_layer("ground", G)  # the reference defines no merge of diverging priors: replicas stay on the shared one
        o.set_layer("groundpatch", C)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{str(k): v for k, v in labels.items()},
             ground=o.layer("ground"), pos=o.position())
    dist.barrier()
    dist.destroy_process_group()


def test_shared_prior_broadcast_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["ground"], r1["ground"]) and np.array_equal(r0["pos"], r1["pos"])
    assert sorted(k for k in r0.files if k.isdigit()) == ["0", "1"] and sorted(k for k in r1.files if k.isdigit()) == ["2", "3"]
    # single-process reference: the owner evaluates all four clouds itself
    dim, res = 33.0, 0.33
    scene = synth.make_scene(seed=5, n_boxes=8, rmin=4.0, rmax=14.0)
    o = Oracle(dim, res)
    o.init_map(0.0, 0.0, 0.0)
    for k in range(3):
        pts, org = synth.lidar_scan(scene, beams=24, az_steps=256, seed=k)
        if k:
            o.update(0.4 * k, 0.0, synth.base_from_map(0.4 * k, 0.0))
        o.filter_cloud(pts, org, 0.0, threads=1)
    for c in range(4):
        pts, org = synth.lidar_scan(scene, ego_xy=(0.8, 0.0), beams=24, az_steps=256, seed=100 + c)
        G, C = o.layer("ground"), o.layer("groundpatch")
        want, _, _ = o.filter_cloud(pts, org, 0.0, threads=1)
        o.set_layer("ground", G)
        o.set_layer("groundpatch", C)
        got = (r0 if c < 2 else r1)[str(c)]
        assert np.array_equal(got, want), f"cloud {c}"
