"""N > 1 on real GPUs (skipped on boxes with fewer than 2): one process per GPU, image strips, the library's own NCCL
gather inside the frame; rank 0's gathered frame must equal the oracle's frame byte for byte."""
import os
import socket

import numpy as np
import pytest

import parity  # noqa: F401
from oracle import harness
from uniformgrid_raytracing_b200 import scenes

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path, transport):
    import torch
    import torch.distributed as dist
    import ugrt  # noqa: F401
    import uniformgrid_raytracing_b200 as ug
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    no_push = transport == "nccl"
    if no_push:
        os.environ["UG_NO_PUSH"] = "1"          # the ncclSend/ncclRecv fallback instead of the IPC push over NVLink
    if transport == "copy":
        os.environ["UG_NO_FUSED_GATHER"] = "1"  # push by a copy-engine node after the frame instead of P2P stores from the shader
    if transport == "fused-noshard":
        os.environ["UG_NO_SHARD_LIGHTS"] = "1"  # every rank builds the light grid in full (the default shards it over the ranks)
    if transport == "fused-nccl-handshake":
        os.environ["UG_NCCL_HANDSHAKE"] = "1"   # ready / done as two ncclAllReduce instead of flags in rank 0's mapped buffer
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # only carries the 128-byte NCCL id
    scene = scenes.scene_a(256, 264)                                # 33 cell rows: uneven strips
    r = ug.Renderer(rank)
    r.set_scene(scene)
    uid = [ug.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    r.comm_init(rank, world, uid[0])
    host = torch.empty(3 * scene.num_pixels, dtype=torch.uint8).pin_memory()
    for frame in range(3):                                          # graph capture, then replays
        r.render_host(r.params(scene, frame=frame, strip=(rank, world)), host.numpy())
    if rank == 0:
        np.save(out_path, host.numpy().copy())
    dist.barrier()
    # strips of unequal height (ug_set_strip_bounds): the gather's offsets follow the boundaries
    nby = scene.height // 8
    bounds = [0] + sorted({1 + (7 * k) % (nby - 2) for k in range(1, world)}) + [nby]
    bounds = bounds if len(bounds) == world + 1 else [nby * k // world for k in range(world + 1)]
    r.set_strip_bounds(bounds, rank)
    for frame in range(2):
        r.render_host(r.params(scene, frame=frame, strip=(rank, world)), host.numpy())
    if rank == 0:
        np.save(out_path + ".bounds.npy", host.numpy().copy())
    dist.barrier()
    # frame-parallel mode: rank k renders frame k of a dynamic scene on its own, rank 0 collects the frames (slot k = rank k)
    dyn = scenes.scene_c(240, 136)
    r.set_scene(dyn)
    for _ in range(2):                                              # capture, replay
        r.render(r.params(dyn, frame=rank, flags=ug.FLAG_GATHER_FRAMES))
    r.synchronize()
    if rank == 0:
        np.save(out_path + ".frames.npy", r.read(ug.BUF_RGB_FULL)[:3 * dyn.num_pixels * world].copy())
    dist.barrier()
    # the same with deferred completion (throughput mode): frame k of every rank lands in slot k & 1 of rank 0's buffer and
    # is whole after the next frame / after comm_flush; 5 steps -> slot 0 holds step 4, slot 1 holds step 3
    fl = ug.FLAG_GATHER_FRAMES | ug.FLAG_GATHER_DEFERRED
    for step in range(5):
        r.render(r.params(dyn, frame=step * world + rank, flags=fl))
    r.comm_flush()
    if rank == 0:
        np.save(out_path + ".deferred.npy", r.read(ug.BUF_RGB_FULL).copy())
    dist.barrier()
    r.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["fused", "copy", "nccl", "fused-noshard", "fused-nccl-handshake"])
def test_strip_gather_over_nccl(oracle, tmp_path, transport):
    """fused: the shading kernels store into rank 0's IPC-mapped frame over NVLink themselves, handshakes are flags in that
    buffer and every rank builds only its share of the light grid's cell rows, stored into every rank's copy (k_lightshare.cu);
    copy: one peer memcpy node after the frame; nccl: grouped ncclSend/ncclRecv; fused-noshard: light grids built in full by
    every rank; fused-nccl-handshake: two ncclAllReduce per frame.  Strips, whole frames, whole frames with deferred completion."""
    no_push = transport == "nccl"
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    scene = scenes.scene_a(256, 264)
    want = harness.run_frame(oracle, scene, threads=0)["rgb"]
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), out, transport), nprocs=world, join=True)
    assert np.array_equal(np.load(out), want)
    assert np.array_equal(np.load(out + ".bounds.npy"), want), "strips of unequal height"
    dyn = scenes.scene_c(240, 136)
    got = np.load(out + ".frames.npy").reshape(world, -1)
    for k in range(world):
        assert np.array_equal(got[k], harness.run_frame(oracle, dyn, frame=k, threads=0)["rgb"]), f"frame of rank {k}"
    both = np.load(out + ".deferred.npy").reshape(2, -1)
    # the ncclSend/ncclRecv fallback has nothing to defer (a receive completes with its data): every frame lands in slot 0
    for slot, step in (((0, 4),) if no_push else ((0, 4), (1, 3))):
        frames = both[slot][:3 * dyn.num_pixels * world].reshape(world, -1)
        for k in range(world):
            want_k = harness.run_frame(oracle, dyn, frame=step * world + k, threads=0)["rgb"]
            assert np.array_equal(frames[k], want_k), f"deferred gather: step {step}, rank {k}"


def _overflow_scene(nlights=0):
    """192 huge triangles below eye level at 1024x1024: the bottom strip's pair count exceeds its initial arena, the top strip's
    does not - only one rank wants to run the frame again."""
    T = 192
    V = np.zeros((T, 3, 3), np.float32)
    for k in range(T):
        y = 20.0 + 0.25 * k
        V[k] = [[-60, y, -60], [60, y, -60], [0, y, 0.5]]
    s = scenes.scene_random(1, 10, 1024, 1024, nlights=nlights)   # no light: a light grid would overflow on both ranks
    s.verts, s.faces, s.face_mat = V.reshape(-1, 3), np.arange(3 * T, dtype=np.int32).reshape(-1, 3), np.zeros(T, np.int32)
    return s


def _worker_overflow(rank, world, port, out_path, nlights=0):
    import torch
    import torch.distributed as dist
    import ugrt  # noqa: F401
    import uniformgrid_raytracing_b200 as ug
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = _overflow_scene(nlights)
    r = ug.Renderer(rank)
    r.set_scene(scene)
    uid = [ug.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    r.comm_init(rank, world, uid[0])
    host = np.zeros(3 * scene.num_pixels, np.uint8)            # pageable on purpose: rank 0 stages the gathered frame itself
    r.render_host(r.params(scene, strip=(rank, world)), host)  # one rank overflows: every rank must run the frame again
    over = [None] * world
    dist.all_gather_object(over, int(r.stats().overflowed))
    r.render_host(r.params(scene, strip=(rank, world)), host)  # and the ranks are still in step afterwards
    if rank == 0:
        np.save(out_path, host)
        np.save(out_path + ".over.npy", np.asarray(over))
    dist.barrier()
    r.close()
    dist.destroy_process_group()


def test_one_rank_overflow_is_a_collective_decision(oracle, tmp_path):
    """the pair-arena overflow retry used to be decided per rank, so a frame in which only one strip overflowed
    left the ranks with different numbers of collectives (hang).  The decision is now a max-allreduce."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    scene = _overflow_scene()
    want = harness.run_frame(oracle, scene, threads=0)["rgb"]
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker_overflow, args=(2, _free_port(), out), nprocs=2, join=True)
    over = np.load(out + ".over.npy")
    assert over.tolist().count(1) == 1, f"exactly one strip was meant to overflow, got {over.tolist()}"
    assert np.array_equal(np.load(out), want)


def test_sharded_light_grid_overflow_falls_back(oracle, tmp_path):
    """With a light, the two ranks build the light grid in row shards (k_lightshare.cu) and a shard of this scene does not fit
    its slot of the shared value list: the frame's collective overflow flag ends the sharding on every rank, the frame is
    rendered again with the light grid built in full (whose arena then has to grow as well) - and still equals the oracle."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    scene = _overflow_scene(1)
    want = harness.run_frame(oracle, scene, threads=0)["rgb"]
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker_overflow, args=(2, _free_port(), out, 1), nprocs=2, join=True)
    assert np.array_equal(np.load(out), want)
