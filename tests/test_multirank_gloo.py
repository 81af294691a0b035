"""world_size-2 (and 3) gloo test of the multi-GPU host logic: work sharding + ONE int64 sum-reduce gives the
same accumulation buffer as a single rank, bit for bit. The per-rank renderer here is the oracle (CPU);
on GPUs the same code path runs with NCCL (bench.py, tests/test_gpu_render.py::test_sharded_render)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, BOX_CAMERA, box_scene


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, spp, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc
    from cudaraytracing_b200.distributed import shard_work, reduce_accum
    verts, mat, obj, mats = box_scene(np.random.default_rng(0), 200)
    S = orc.Scene().add_arrays(verts, mat, obj, mats)
    S.build_new_bvh(2)
    W, H = 40, 30
    M = orc.inverse_view_matrix(BOX_CAMERA["eye"], BOX_CAMERA["lookat"], BOX_CAMERA["up"])
    b, e = shard_work(W * H, spp, rank, world)
    assert b % (W * H) == 0 and e % (W * H) == 0           # spp >= world: whole-sample shards
    acc, _ = S.render(BOX_CAMERA["eye"], M, BOX_CAMERA["fovy"], W, H, b // (W * H), e // (W * H), 0.6, 2, seed=3, threads=1)
    t = torch.from_numpy(acc)
    reduce_accum(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,spp", [(2, 4), (3, 5)])
def test_sharded_reduce_is_bit_exact(orc, tmp_path, world, spp):
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(world, _free_port(), spp, out), nprocs=world, join=True)
    verts, mat, obj, mats = box_scene(np.random.default_rng(0), 200)
    S = orc.Scene().add_arrays(verts, mat, obj, mats)
    S.build_new_bvh(2)
    M = orc.inverse_view_matrix(BOX_CAMERA["eye"], BOX_CAMERA["lookat"], BOX_CAMERA["up"])
    full, _ = S.render(BOX_CAMERA["eye"], M, BOX_CAMERA["fovy"], 40, 30, 0, spp, 0.6, 2, seed=3)
    assert np.array_equal(np.load(out), full)
    from cudaraytracing_b200.distributed import resolve_numpy
    lin, rgb = resolve_numpy(full, spp)
    olin, orgb = orc.resolve(full, 40 * 30, spp)
    assert np.array_equal(lin, olin) and np.abs(rgb.astype(int) - orgb.astype(int)).max() <= 1
