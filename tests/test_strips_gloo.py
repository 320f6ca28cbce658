"""N > 1 host logic on CPU: two / four gloo ranks each own a strip of cell rows, gather to rank 0, reassemble; the result
must be the single-rank image byte for byte.  (The per-strip rendering itself is CUDA-only and is covered by
tests/test_gpu_parity.py::test_strips_equal_full_frame; here the strips are cut from the oracle's frame.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity  # noqa: F401
from oracle import harness
from uniformgrid_raytracing_b200 import scenes
import strips_helpers as strips


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, width, height, image, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = strips.strip_nbytes(width, height, world)
    start = sum(sizes[:rank])
    strip = torch.zeros(max(sizes), dtype=torch.uint8)
    strip[:sizes[rank]] = torch.from_numpy(image[start:start + sizes[rank]].copy())   # this rank's rows of the frame
    gathered = strips.gather_strips(strip, width, height, rank, world)
    if rank == 0:
        full = strips.assemble(gathered, width, height, world)
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gather_reassembles_frame(oracle, tmp_path, world):
    scene = scenes.scene_random(21, 400, 64, 72, nlights=1)          # 9 cell rows: uneven strips (4 + 5; 2 + 2 + 2 + 3)
    image = harness.run_frame(oracle, scene, threads=0)["rgb"]
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), scene.width, scene.height, image, out), nprocs=world, join=True)
    assert np.array_equal(np.load(out), image)
