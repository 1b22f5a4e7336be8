"""N > 1 host logic on CPU: two `gloo` ranks shard a batch by cloud, run the chain on their shard (with
the CPU oracle standing in for the kernels -- the checker, not the product) and the gathered result must
equal the single-process result bit for bit; timing is reduced as the max over ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from torch_points3d_b200 import sharding, synthetic


def test_cloud_range_partitions_exactly():
    for n in (0, 1, 7, 8, 16, 17):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharding.cloud_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.cloud_range(4, 2, 2)


def test_shard_batched_offsets_and_renumbering():
    sizes = [3, 1, 5, 2, 4]
    batch = torch.repeat_interleave(torch.arange(5), torch.tensor(sizes))
    pos = torch.arange(batch.numel() * 3, dtype=torch.float32).view(-1, 3)
    feat = torch.arange(batch.numel(), dtype=torch.float32).view(-1, 1)
    rows = 0
    for rank in range(2):
        p, b, off, f = sharding.shard_batched(pos, batch, 2, rank, feat)
        assert off == rows and b.min().item() == 0
        assert torch.equal(p, pos[off:off + p.shape[0]]) and torch.equal(f, feat[off:off + p.shape[0]])
        rows += p.shape[0]
    assert rows == batch.numel()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _chain(xyz, feat):
    """FPS -> ball query -> grouping on a [b, N, 3] shard (oracle = stand-in for the kernels on CPU)"""
    idx = oracle.fps_dense(xyz, 64)
    new_xyz = np.take_along_axis(xyz, idx[..., None], 1)
    ridx, _ = oracle.ball_query_dense(0.3, 16, xyz, new_xyz)
    return idx, ridx, oracle.group_fwd(feat, ridx)


def _worker(rank, world, port, xyz, feat, pos, batch, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # dense layout
        mine = sharding.shard_dense(torch.from_numpy(xyz), world, rank).numpy()
        fmine = sharding.shard_dense(torch.from_numpy(feat), world, rank).numpy()
        idx, ridx, grouped = _chain(np.ascontiguousarray(mine), np.ascontiguousarray(fmine))
        parts = [None] * world
        dist.all_gather_object(parts, (idx, ridx, grouped))
        # batched layout: knn inside each rank's clouds, indices shifted back to global rows
        p, b, off = sharding.shard_batched(torch.from_numpy(pos), torch.from_numpy(batch), world, rank)
        ptr = oracle.batch_to_ptr(b.numpy())
        col, _ = oracle.knn_csr(p.numpy(), p.numpy(), 4, ptr, ptr)
        cols = [None] * world
        dist.all_gather_object(cols, np.where(col >= 0, col + off, col))
        slowest = sharding.max_over_ranks(10.0 + rank)
        if rank == 0:
            out.put((parts, cols, slowest))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process():
    world = 2
    xyz = synthetic.surface_clouds(5, 300, seed=77)           # 5 clouds: uneven 3 + 2 split
    feat = np.random.default_rng(77).standard_normal((5, 4, 300)).astype(np.float32)
    pos, batch = synthetic.indoor_spheres(num_spheres=3, raw_points=1500, radius=0.3, seed=78)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, xyz, feat, pos, batch, out)) for r in range(world)]
    for p in procs:
        p.start()
    parts, cols, slowest = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.concatenate([a[0] for a in parts])
    ridx = np.concatenate([a[1] for a in parts])
    grouped = np.concatenate([a[2] for a in parts])
    ref_idx, ref_ridx, ref_grouped = _chain(xyz, feat)
    assert np.array_equal(idx, ref_idx) and np.array_equal(ridx, ref_ridx) and np.array_equal(grouped, ref_grouped)
    ptr = oracle.batch_to_ptr(batch)
    ref_col, _ = oracle.knn_csr(pos, pos, 4, ptr, ptr)
    assert np.array_equal(np.concatenate(cols), ref_col)
    assert slowest == 11.0  # max over ranks


def test_reference_arm_under_torchrun_prints_one_line():
    """bench.py --impl reference launched like the driver launches it for N > 1: rank 0 alone measures and prints
    ONE JSON line carrying the same metric / unit / workload as the b200 arm; the other rank exits 0 silently"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(root, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    sys.path.insert(0, root)
    import bench
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["config"]["workload"] == bench.WORKLOAD and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
