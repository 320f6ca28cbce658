#!/usr/bin/env python
"""bench.py -- Mrays/s and ms/frame INCLUDING the per-frame grid rebuild (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W [--config B] [--impl reference]

One "step" = one frame of the hot path: [vertex transform ->] camera-grid rebuild -> primary rays -> per light
(light-grid rebuild, shadow-ray regroup, shadow rays) -> shading to RGB8.  rays/frame = pixels * (1 + lights)
(the reference shoots a shadow ray for every pixel, misses included).

  value     device-resident frames through the C-ABI (ug_render_frame), timed per frame with CUDA events on the
            stream the library launches on; L2 is flushed (256 MiB write) between timed frames.
  e2e       the same frames through the host-buffer C-ABI call (ug_render_frame_host): pinned H2D of the frame
            constants and D2H of the RGB8 image inside the timed region, wall clock.
  N = 1     config B (the configuration BASELINE.json quotes the metric on); the line also carries config D's one-GPU frame
            time under "config_D_single_gpu", the denominator of the N > 1 series.
  N > 1     one process per GPU (torchrun), scene replicated.  Default: config D (282 000 triangles, 4096x4096) as ONE frame
            cut into N image strips of cell rows - the tile sharding north-star names, strong scaling of the frame latency.
            Every rank builds the camera grid of its own sub-frustum, traces and shades its strip, and the shading kernel
            stores the RGB8 pixels straight into rank 0's IPC-mapped frame over NVLink (P2P stores; NCCL carries the two
            4-byte handshakes), all inside the frame's CUDA graph.  After the timed region rank 0 renders the same frame alone
            and the gathered frame must hash to the same bytes ("gathered_frame_check"); that one-GPU time is reported as
            "single_gpu_same_frame".  The frame-parallel mode (one whole frame per GPU per step, weak scaling, deferred
            gather) is measured beside it and reported under "frames"; --shard frames makes it the headline.
  --light-grid spherical --shader spot
            the variant the committed display() runs (main.cu:186-219) instead of the one north-star names.
  roofline  the dominant kernel (timed live with one CUDA event per stage): algorithmic bytes/launch vs the
            measured HBM peak, plus the FP32-issue roofline that actually bounds ray traversal.
  cpu_baseline / --impl reference
            the reference's own kernel code compiled as host C++ (oracle/_ref/libugref.so) on the box's host cores.

There is no CPU fallback on the product arm: without a GPU or without libugrt.so this script fails.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s incl. per-frame grid rebuild"
FP32_ISSUE_PER_TEST = 48  # FP32 instructions of one Moeller-Trumbore test without FMA


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured", float(p.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_scene(name: str):
    import ugrt  # noqa: F401
    from uniformgrid_raytracing_b200 import scenes
    if name == "A":
        return scenes.scene_a()
    s = scenes.SCENES[name]()
    return s


def workload_name(scene, world, bounces=0, mode="single"):
    return (f"config {scene.name}: {scene.num_tris} triangles, {scene.width}x{scene.height}, primary + shadow"
            + (f" + {bounces} reflection bounce(s)" if bounces else "") + " rays, "
            f"{len(scene.lights)} light(s), full grid rebuild every frame"
            + (", per-frame vertex transform" if scene.dyn_size else "")
            + (f", {world} image strips" if mode == "strips" else "") + (f", {world} frames per step (one per GPU)" if mode == "frames" else ""))


# --------------------------------------------------------------------------------------------- reference arm
def cpu_reference(scene, frames: int, warmup: int, budget_s: float = 0.0):
    """Time the reference's own CPU code (host-compiled kernels, oracle/_ref) - falls back to the C port.
    budget_s > 0: run whole frames until about that much CPU time is spent (at least `frames`)."""
    from oracle import harness
    kind = "ref" if harness.available("ref") else "oracle"
    if kind == "oracle" and not harness.available("oracle"):
        harness.build("oracle")
    lib = harness.load(kind)
    cores = os.cpu_count() or 1
    for _ in range(max(warmup, 0)):
        harness.run_frame(lib, scene, threads=cores)
    t0 = time.perf_counter()
    done = 0
    while done < frames or (budget_s > 0 and time.perf_counter() - t0 < budget_s and done < 1000):
        harness.run_frame(lib, scene, threads=cores)
        done += 1
    frames = done
    dt = time.perf_counter() - t0
    rays = scene.rays_per_frame() * frames
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": lib.get_threads(), "kind": "reference" if kind == "ref" else "port",
            "sample": f"{frames} full frame(s) of the same workload ({scene.rays_per_frame()} rays each), {dt:.2f} s",
            "ms_per_frame": dt / frames * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene = make_scene(args.config or ("B" if args.gpus <= 1 else "D"))   # the product arm's workload at this N
    # each step = one full frame of the workload; K is honoured unless that would take more than ~2 minutes (the first
    # frame is timed to find out), in which case fewer steps are run and the line says how many
    from oracle import harness
    kind = "ref" if harness.available("ref") else "oracle"
    if kind == "oracle" and not harness.available("oracle"):
        harness.build("oracle")
    t0 = time.perf_counter()
    harness.run_frame(harness.load(kind), scene, threads=os.cpu_count() or 1)
    one = time.perf_counter() - t0
    frames = max(1, min(args.steps, int(120.0 / max(one, 1e-3))))
    base = cpu_reference(scene, frames, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": frames,
            "warmup": min(args.warmup, 1), "ms_per_step": base["ms_per_frame"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(scene, 1), "note": "reference kernel sources compiled as host C++ (oracle/_ref), OpenMP over "
                       "cells/pixels on all host cores; each step = one full frame; the reference has no reflection code "
                       "(SURVEY.md 0.3), so its frames hold the primary and shadow rays only",
                       "steps_requested": args.steps},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------- product arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ugrt  # noqa: F401
    import uniformgrid_raytracing_b200 as ug

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path (use --impl reference for the CPU arm)")
    # stdout carries exactly one JSON line: everything libraries print (NCCL banner ...) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_kind, sm_max = load_peaks()

    config = args.config or ("B" if world == 1 else "D")
    scene = make_scene(config)
    r = ug.Renderer(local)
    r.set_scene(scene)
    # the library launches on the stream it is given (NULL would mean its own stream): a dedicated torch stream, made
    # current, so torch.cuda.Event.record() brackets exactly the library's work
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    r.set_stream(work_stream.cuda_stream)
    # N > 1, two ways to shard:
    #   frames (default): one whole frame per GPU per step, the frames gathered to rank 0 over NCCL - per-GPU work is fixed,
    #                     so this is the weak-scaling (throughput) series; N = 1 is the same workload
    #   strips          : ONE frame cut into image strips of cell rows, gathered inside the frame graph - strong scaling
    #                     (latency) of a single frame; measured after the main loop too and reported under "strips"
    mode = "single" if world == 1 else ("frames" if args.shard == "frames" else "strips")
    nby = scene.height // 8

    def strip_of(m):
        return (rank, world) if m == "strips" else (0, 1)

    def join_comm():   # the library gathers the strips itself (P2P stores over NVLink, handshakes as flags in rank 0's mapped buffer, inside the frame's CUDA graph; NCCL bootstraps the IPC handles)
        uid = [ug.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        r.comm_init(rank, world, uid[0])

    bounces = scene.bounces if args.bounces < 0 else args.bounces
    if not any(m.reflect > 0 for m in scene.materials):
        bounces = 0
    base_flags = (ug.FLAG_SPHERICAL_LIGHTS if args.light_grid == "spherical" else 0) | (ug.FLAG_SPOT_SHADE if args.shader == "spot" else 0)

    # ---- strips: move the strip boundaries until every rank's strip costs the same (measured, before the communicator exists)
    bounds, balance = None, None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)            # > 126 MB L2
    def rebalance(T, bounds):
        """Equal-cost strip boundaries from the measured strip times T.  A strip costs F + (its rows' cost): F, the part that does
        not shrink with the strip (launch chain, light grid, projecting every triangle), is taken as 0.3 x the fastest strip -
        without it the correction undershoots and the loop converges slowly.  The last step of a loop is not kept blindly: the
        best boundaries seen (smallest slowest strip) are."""
        F = 0.3 * float(T.min())
        dens = np.concatenate([np.full(bounds[k + 1] - bounds[k], max(T[k] - F, 1e-6) / (bounds[k + 1] - bounds[k])) for k in range(world)])
        cum = np.concatenate([[0.0], np.cumsum(dens)])
        target = [int(np.searchsorted(cum, cum[-1] * k / world)) for k in range(1, world)]
        new = [0] + [int(round(0.75 * t + 0.25 * b)) for t, b in zip(target, bounds[1:-1])] + [nby]
        for k in range(1, world + 1):
            new[k] = max(new[k], new[k - 1] + 1)
        for k in range(world - 1, 0, -1):
            new[k] = min(new[k], new[k + 1] - 1)
        return new

    prime_img = torch.empty(3 * scene.num_pixels, dtype=torch.uint8).pin_memory()
    invalid = []

    def prime(strip, flags=0):
        """ug_render_frame is asynchronous: a pair arena that proves too small only shows in ug_get_stats.  Every series of timed
        frames is therefore preceded by one frame through ug_render_frame_host, which grows the arenas and renders the frame
        again until it fits (collective when the frame is gathered) - config E needs that, A-D never overflow."""
        r.render_host(r.params(scene, frame=0, flags=base_flags | flags, strip=strip, bounces=bounces, slabs=args.slabs), prime_img.numpy())

    def check_valid(what):
        """after a series: the last frame must not have overflowed an arena (a frame that did is incomplete and fast)"""
        if r.stats().overflowed:
            invalid.append(what)

    def measure_strips(own: bool):
        """median time of this rank's strip over 9 flushed frames, gathered from every rank"""
        pcal = r.params(scene, frame=0, flags=base_flags, strip=(rank, world), bounces=bounces, slabs=args.slabs)
        for _ in range(3):
            r.render(pcal)
        ts = []
        for _ in range(9):
            if not args.no_flush:
                flush.zero_()                      # the timed frames start from a flushed L2 as well
            r.render(pcal)
            st_ = r.stats()
            ts.append(st_.ms_own if own else st_.ms_frame)
        mine = torch.tensor([float(np.median(ts))], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        return np.asarray([float(x.item()) for x in allt])

    balancing = world > 1 and not args.equal_strips and mode == "strips"
    history = []
    if balancing:   # first without the communicator: every rank renders its strip on its own
        bounds = [nby * k // world for k in range(world + 1)]
        r.set_screen(scene.width, scene.height, scene.fovy)
        for it in range(4):
            r.set_strip_bounds(bounds, rank)
            prime((rank, world))
            T = measure_strips(own=False)
            history.append({"bounds": list(bounds), "ms": [round(float(x), 4) for x in T], "gathered": False})
            if T.max() / T.mean() < 1.015 or it == 3:
                break
            bounds = rebalance(T, bounds)
        r.set_strip_bounds(bounds, rank)
    lo, hi = ((bounds[rank], bounds[rank + 1]) if bounds else (nby * rank // world, nby * (rank + 1) // world)) if mode == "strips" else (0, nby)
    n_local = scene.width * (hi - lo) * 8
    if world > 1:
        join_comm()
    if balancing:   # then inside the real gathered frames, on each rank's own time (ug_stats.ms_own: first kernel -> its last remote store)
        best = None
        prime((rank, world))
        for it in range(6):
            T = measure_strips(own=True)
            history.append({"bounds": list(bounds), "ms": [round(float(x), 4) for x in T], "gathered": True})
            if best is None or T.max() < best[0]:
                best = (float(T.max()), list(bounds))
            if T.max() / T.mean() < 1.015 or it == 5:
                break
            bounds = rebalance(T, bounds)
            r.set_strip_bounds(bounds, rank)
            prime((rank, world))
        if bounds != best[1]:
            bounds = best[1]
            r.set_strip_bounds(bounds, rank)
            prime((rank, world))
        balance = {"row_bounds": list(bounds), "iterations": history,
                   "note": "strip boundaries (cell rows) moved until the strips' measured times agree: image-space cost is not uniform"}
        lo, hi = bounds[rank], bounds[rank + 1]
        n_local = scene.width * (hi - lo) * 8
    host_img = torch.empty(3 * scene.num_pixels, dtype=torch.uint8).pin_memory()
    rays_frame = scene.rays_per_frame()

    def step(i, flags=0, m=None):
        m = m or mode
        flags |= base_flags
        if m == "frames":   # whole frames, collected on rank 0 by the library itself inside the frame graph
            # deferred completion: frame i is whole on rank 0 once frame i+1 has passed its "ready" handshake, or after
            # comm_flush() - the per-frame "done" handshake is off the critical path, the timed region ends with the flush
            r.render(r.params(scene, frame=i * world + rank, flags=flags | ug.FLAG_GATHER_FRAMES | ug.FLAG_GATHER_DEFERRED, bounces=bounces, slabs=args.slabs))
        else:
            r.render(r.params(scene, frame=i, flags=flags, strip=strip_of(m), bounces=bounces, slabs=args.slabs))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also captures the CUDA graph and grows the pair arena if the scene needs it)
    r.render_host(r.params(scene, frame=0, flags=base_flags, strip=strip_of(mode), bounces=bounces, slabs=args.slabs), host_img.numpy())
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()

    # ---- timed: K frames, one CUDA-event pair per frame, L2 flushed between frames
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    if world > 1:
        # the ranks leave the host barrier up to a millisecond apart, and the first frame's handshake would charge that skew to
        # frame 0: one more untimed frame lines the GPUs up (its "done" handshake ends on every rank at the same time)
        step(max(args.warmup, 3))
    wall0 = time.perf_counter()
    for i in range(args.steps):
        if not args.no_flush:
            flush.zero_()
        # N > 1: no host barrier per step - the gather inside the frame (ready / done handshakes) keeps the ranks in
        # lockstep, and a host barrier here only adds its own wake-up skew to every frame
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    flush_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    flush_ev[0].record()
    if mode == "frames":
        r.comm_flush()           # the last frame's gather completes here: part of the timed region
    flush_ev[1].record()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    check_valid("timed frames")
    # the same K frames with a warm L2 (what a render loop sees between consecutive frames), reported next to the flushed figure
    warm = []
    for i in range(min(args.steps, 20)):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(); step(i); b_.record()
        warm.append((a_, b_))
    torch.cuda.synchronize()
    warm_t = torch.tensor([a_.elapsed_time(b_) for a_, b_ in warm], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(warm_t, op=dist.ReduceOp.MAX)
    warm_ms = float(warm_t.mean().item())
    per = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(per, op=dist.ReduceOp.MAX)   # a frame is done when its slowest strip (and the gather) is
    per = per.cpu().numpy()
    tail = torch.tensor([flush_ev[0].elapsed_time(flush_ev[1])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tail, op=dist.ReduceOp.MAX)
    total_ms = float(per.sum()) + (float(tail.item()) if mode == "frames" else 0.0)
    st = r.stats()
    launches_per_frame = st.kernel_launches
    refl = torch.tensor([float(st.reflect_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(refl)                        # every rank traces the reflection rays of its own strip
    frames_per_step = world if mode == "frames" else 1
    rays_frame = scene.rays_per_frame() * frames_per_step + int(refl.item())   # rays of one step

    # ---- e2e: host buffers in, host image out, wall clock.  Every step: the frame constants are read from pinned host memory
    #      by the frame's first kernel (H2D), the kernels run [+ strip gather], the RGB8 image is copied to pinned host memory
    #      (D2H; strips: the gathered frame on rank 0; frames: every rank reads its own frame back over its own PCIe link).
    #      Two ways through the C-ABI: one frame at a time (ug_render_frame_host, returns with the image) and with two frames
    #      in flight (ug_render_frame_host_async + ug_wait_frame: frame i+1 renders while frame i's image crosses PCIe).
    e2e_steps = max(3, min(args.steps, 50))
    out_np = host_img.numpy()[:3 * n_local]          # pinned: the copy engine writes it directly
    host_img2 = torch.empty(3 * scene.num_pixels, dtype=torch.uint8).pin_memory()
    bufs = [host_img.numpy() if mode == "strips" else out_np, host_img2.numpy() if mode == "strips" else host_img2.numpy()[:3 * n_local]]

    def e2e_params(i):
        return r.params(scene, frame=i * frames_per_step + (rank if mode == "frames" else 0), flags=base_flags, strip=strip_of(mode),
                        bounces=bounces, slabs=args.slabs)

    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        r.render_host(e2e_params(i), bufs[0])
    barrier()
    e2e_sync_dt = time.perf_counter() - t0
    r.render_host_async(e2e_params(0), bufs[0]); r.wait_frame(0)     # the graph with the read-back's wait node is captured outside the timed region
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        r.render_host_async(e2e_params(i), bufs[i & 1])
        if i >= 1:
            r.wait_frame(1)                           # frame i-1's image is in host memory: its buffer can be consumed and handed out again
    r.wait_frame(0)
    barrier()
    e2e_dt = time.perf_counter() - t0
    check_valid("e2e frames")
    e2e_t = torch.tensor([e2e_dt, e2e_sync_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_dt, e2e_sync_dt = float(e2e_t[0].item()), float(e2e_t[1].item())

    # ---- per-stage timing (one event per stage, no graph) -> dominant kernel + rooflines
    stage = {}
    reps = 5
    for i in range(reps):
        r.render(r.params(scene, frame=i, flags=ug.FLAG_STAGE_EVENTS | base_flags, strip=strip_of(mode), bounces=bounces, slabs=args.slabs))
        for k, v in r.stage_times().items():
            stage[k] = stage.get(k, 0.0) + v / reps
    st2 = r.stats()

    # ---- N > 1: the same workload once more in the OTHER sharding mode (strips <-> frames), reported beside the headline
    other = None
    if world > 1:
        om = "frames" if mode == "strips" else "strips"
        if om == "frames":
            prime((0, 1), ug.FLAG_GATHER_FRAMES)
        else:
            prime((rank, world))
        for i in range(5):
            step(i, m=om)
        if om == "frames":
            r.comm_flush()
        barrier()
        k2 = min(args.steps, 30)
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k2)]
        for i in range(k2):
            if not args.no_flush:
                flush.zero_()
            ev2[i][0].record()
            step(i, m=om)
            ev2[i][1].record()
        tail2 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        tail2[0].record()
        if om == "frames":
            r.comm_flush()
        tail2[1].record()
        barrier()
        t2 = torch.tensor([a.elapsed_time(b) for a, b in ev2] + [tail2[0].elapsed_time(tail2[1])], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        check_valid("frames of the other sharding mode")
        ms2 = float(t2.sum().item()) / k2
        fps2 = world if om == "frames" else 1
        other = {"mode": om, "ms_per_step": ms2, "steps": k2, "scaling": "weak" if om == "frames" else "strong", "frames_per_step": fps2,
                 "note": ("one whole frame per GPU per step, gathered to rank 0 with deferred completion" if om == "frames" else
                          "one frame cut into %d image strips of cell rows" % world) + ("; shading kernels store into rank 0's frame over NVLink (P2P); ready = ncclAllReduce that doubles as done of the frame before" if om == "frames" else
                             "; shading kernels store into rank 0's frame over NVLink (P2P), ready / done handshakes as flags in rank 0's mapped buffer, light grids built in row shards")}

    # ---- N > 1, strips: rank 0's gathered frame must be the frame one GPU renders alone; that render is also the 1-GPU time of this workload
    check, single = None, None
    if world > 1:
        import hashlib
        if bounds:
            r.set_strip_bounds(bounds, rank)
        r.render_host(r.params(scene, frame=0, flags=base_flags, strip=(rank, world), bounces=bounces, slabs=args.slabs), host_img.numpy())
        barrier()
        if rank == 0:
            h_gather = hashlib.sha256(host_img.numpy().tobytes()).hexdigest()
            p1 = r.params(scene, frame=0, flags=base_flags, strip=(0, 1), bounces=bounces, slabs=args.slabs)
            prime((0, 1))        # not gathered: rank 0 alone
            for _ in range(3):
                r.render(p1)
            r.synchronize()
            h_single = hashlib.sha256(r.read(ug.BUF_RGB).tobytes()).hexdigest()
            evs = []
            for i in range(min(args.steps, 20)):
                if not args.no_flush:
                    flush.zero_()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record(); r.render(p1); b_.record()
                evs.append((a_, b_))
            torch.cuda.synchronize()
            ms1 = float(np.mean([a_.elapsed_time(b_) for a_, b_ in evs]))
            check_valid("single-GPU frame")
            check = {"sha256_gathered": h_gather, "sha256_single_gpu": h_single, "equal": h_gather == h_single}
            single = {"ms_per_frame": ms1, "value": scene.rays_per_frame() / (ms1 * 1e-3) / 1e6, "unit": "Mrays/s", "steps": len(evs),
                      "note": "rank 0 renders the same frame alone (strip 0 of 1, no gather) in the same process, same timing rules"}
        barrier()
        r.render(r.params(scene, frame=0, flags=base_flags, strip=strip_of(mode), bounces=bounces, slabs=args.slabs))   # back to the run's geometry
        r.synchronize()

    # ---- N = 1: config D's one-GPU frame time (the workload of the N > 1 series), so the scaling series has its own denominator
    config_d = None
    if world == 1 and config == "B" and not args.no_config_d:
        sd = make_scene("D")
        rd = ug.Renderer(local)
        rd.set_scene(sd)
        rd.set_stream(work_stream.cuda_stream)
        pd_ = rd.params(sd, frame=0, flags=base_flags)
        rd.render_host(pd_, torch.empty(3 * sd.num_pixels, dtype=torch.uint8).pin_memory().numpy())   # grows the arenas if the scene needs it
        for _ in range(4):
            rd.render(pd_)
        torch.cuda.synchronize()
        evs = []
        for i in range(20):
            if not args.no_flush:
                flush.zero_()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(); rd.render(pd_); b_.record()
            evs.append((a_, b_))
        torch.cuda.synchronize()
        msd = float(np.mean([a_.elapsed_time(b_) for a_, b_ in evs]))
        config_d = {"workload": workload_name(sd, 1), "ms_per_frame": msd, "value": sd.rays_per_frame() / (msd * 1e-3) / 1e6, "unit": "Mrays/s", "steps": 20}
        rd.close()

    # ---- UG_TIMELINE=1: marker times of every rank's frame (development aid; the markers are extra one-thread kernels, so the
    #      env variable is never set for a bench line that is reported)
    timeline = None
    if os.environ.get("UG_TIMELINE"):
        import ctypes as C
        L_ = ug.lib()
        L_.ug_timeline.restype = C.c_int
        L_.ug_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        acc = []
        for it in range(8):
            flush.zero_()
            step(it)
            r.synchronize()
            out = (C.c_double * 64)()
            L_.ug_timeline(r.ctx, out)
            acc.append(np.array(out[:]))
        tmed = torch.tensor(np.median(np.stack(acc[2:]), axis=0) / 1000.0, dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(tmed) for _ in range(world)]
        if world > 1:
            dist.all_gather(allt, tmed)
        else:
            allt = [tmed]
        names = {1: "cam_grid", 11: "light_grid", 6: "ready", 2: "primary", 13: "regroup", 14: "shadow", 3: "shade", 5: "gather"}
        timeline = {str(k): {n: round(float(allt[k][i].item()), 1) for i, n in names.items()} for k in range(world)}

    invalid_all = list(invalid)
    if world > 1:
        got = [None] * world
        dist.all_gather_object(got, invalid)
        invalid_all = sorted({f"rank {k}: {w}" for k, lst in enumerate(got) for w in lst})
    if rank == 0:
        try:
            issue_measured = r.fp32_issue_peak() / 1e12
        except Exception:
            issue_measured = None
        P, N, Cc = st2.cam_pairs, st2.num_pixels, st2.num_cells
        T = scene.num_tris
        # single-kernel stages only (grid builds / regroup are several kernels); algorithmic bytes per launch from SURVEY.md 8(d)
        R1 = int(refl.item()) // max(frames_per_step, 1)      # reflection rays of one frame (all bounces)
        Pu, Cu = int(st2.reflect_pairs), int(2.0 * T)         # uniform grid: pairs, cells (cells_per_tri = 2)
        kern = {"trace": ("k_trace", 8 * Cc + 52 * P + 8 * N, 64 * P, "FP32-issue bound (ray-triangle tests), not HBM bound: see fp32_issue"),
                "shade": ("k_shade", (8 + len(scene.lights) / 8.0) * N + 3 * N, None, "HBM streaming; SURVEY 8(d): (8 + L/8) N in + 3 N out"),
                "reflect_rays": ("k_refl_trace", 8 * R1, None,
                                 "SURVEY 8(d): compaction traffic 8 R_b; the 3-D DDA itself is latency / issue bound (dependent cell and record "
                                 "fetches, divergent rays) - see `issue` (lanes active per warp instruction x issue-slot utilisation, from the "
                                 "committed ncu capture); the stage also holds the ray-spawn kernel (~10 us)")}
        for l in range(len(scene.lights)):
            kern[f"light{l}_shadow"] = ("k_shadow", None, None, "")

        def roof(stage_name):
            name, alg_bytes, tests, note = kern[stage_name]
            t_s = stage[stage_name] * 1e-3
            achieved = alg_bytes / t_s / 1e9
            traffic, issue = None, None   # from the committed ncu capture (profiles/traffic.json), not measured in this run
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    prof = json.load(f)
                traffic, issue = prof.get(name), prof.get("issue_active_pct", {}).get(name)
            except Exception:
                pass
            out = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                   "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full capture of the same command, committed)" if traffic is not None else None,
                   "peak_source": peak_kind, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": stage[stage_name], "note": note}
            if issue is not None:   # from the committed ncu capture of the same command (profiles/), not measured live
                out["sm_issue_active_pct_ncu"] = issue
                lanes = prof.get("lanes_active", {}).get(name)
                if lanes is not None:
                    out["issue"] = {"issue_active_pct": issue, "lanes_active_of_32": lanes, "useful_issue_frac": issue / 100.0 * lanes / 32.0,
                                    "source": "profiles/ (ncu --set full of the same command), not measured in this run"}
            if tests:
                nominal = 148 * 128 * sm_max * 1e6 / 1e12   # T lane-ops/s at the max SM clock
                issue_peak = issue_measured if issue_measured else nominal   # FMUL + FADD micro-benchmark of this run (k_peak.cu)
                ach = tests * FP32_ISSUE_PER_TEST / t_s / 1e12
                out["fp32_issue"] = {"achieved": ach, "peak": issue_peak, "unit": "T lane-ops/s", "frac": ach / issue_peak,
                                     "peak_source": "measured in this run (ug_fp32_issue_peak: FMUL + FADD, all SMs)" if issue_measured else "nominal",
                                     "peak_nominal": nominal, "frac_of_nominal": ach / nominal,
                                     "ray_triangle_tests_per_launch": tests, "fp32_instr_per_test": FP32_ISSUE_PER_TEST}
            return out

        dom = max((k for k in kern if k in stage and stage[k] > 0), key=lambda k: stage[k])
        if kern[dom][1] is None:   # shadow traversal: report the primary-traversal kernel instead (same inner loop, exact counts)
            dom = "trace"
        roofline = roof(dom)
        roofline_all = {kern[k][0]: roof(k) for k in kern if k in stage and stage[k] > 0 and kern[k][1] is not None}
        frame_bytes = 0
        for nm, Pg, Cg in [("cam", P, Cc)] + [(f"l{l}", st2.light_pairs[l], (scene.width // 8) * (scene.height // 8)) for l in range(len(scene.lights))]:
            frame_bytes += 48 * T + 4 * Pg + 8 * Cg
        frame_bytes += 8 * Cc + 52 * P + 8 * N + len(scene.lights) * 16 * N + int((8 + len(scene.lights) / 8.0) * N) + 3 * N + 8 * R1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_reference(scene, 3, 1, budget_s=10.0)
                cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the checker is optional for the bench line
                cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        line = {
            "metric": METRIC, "value": rays_frame * args.steps / (total_ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if mode == "strips" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(scene, world, bounces, mode), "rays_per_step": rays_frame, "frames_per_step": frames_per_step,
                       "parallelism": {"single": "one GPU", "frames": f"{world} GPUs, one frame each per step, frames stored into rank 0's buffer over NVLink by the shading kernels (P2P; one ncclAllReduce per frame as ready + deferred done)",
                                       "strips": f"{world} GPUs, one frame in {world} image strips of cell rows, stored into rank 0's frame over NVLink by the shading kernels (P2P; ready / done handshakes as flags in rank 0's IPC-mapped buffer; each light grid built in row shards by the ranks and stored into every rank's copy over NVLink; all inside the frame graph)"}[mode],
                       "rays": {"primary": scene.num_pixels * frames_per_step, "shadow": scene.num_pixels * len(scene.lights) * frames_per_step,
                                "reflection": int(refl.item())},
                       "l2": "flushed between timed frames (256 MiB write)" if not args.no_flush else "not flushed",
                       "timing": "CUDA events per frame on the launching stream, max over ranks, summed over the K frames",
                       "pairs_camera_grid": int(P), "pairs_light_grids": [int(st2.light_pairs[l]) for l in range(len(scene.lights))],
                       "light_grid": args.light_grid, "shader": args.shader, "num_slabs": args.slabs},
            "ms_per_frame_warm_l2": warm_ms, "ms_per_frame_p50": float(np.median(per)), "ms_per_frame_p10": float(np.percentile(per, 10)),
            "ms_per_frame_p90": float(np.percentile(per, 90)), "ms_per_frame_all": [round(float(x), 4) for x in per], "wall_s_timed_region": wall,
            "e2e": {"value": rays_frame * e2e_steps / e2e_dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(ug.lib().ug_frame_const_bytes()) * world,
                    "d2h_bytes_per_step": 3 * scene.num_pixels * frames_per_step, "steps": e2e_steps, "ms_per_step": e2e_dt / e2e_steps * 1e3,
                    "frames_in_flight": 2, "api": "ug_render_frame_host_async + ug_wait_frame: host constants in, host image out every step, the image "
                    "copy of frame i overlaps the rendering of frame i+1; wall clock over all steps, barrier + synchronize on both sides",
                    "one_frame_at_a_time": {"value": rays_frame * e2e_steps / e2e_sync_dt / 1e6, "ms_per_step": e2e_sync_dt / e2e_steps * 1e3,
                                            "api": "ug_render_frame_host (returns with the image in host memory)"}},
            "gpu_launches": int(launches_per_frame) * args.steps * world,
            "kernels_per_frame": int(launches_per_frame),
            "library": {"path": os.path.relpath(ug.LIB_PATH, ROOT), "version": ug.lib().ug_version().decode(),
                        "env": {k: v for k, v in os.environ.items() if k.startswith("UG_") or k == "UGRT_LIB"}},
            "stage_ms": {k: round(v, 4) for k, v in stage.items()},
            "frame_algorithmic_bytes": frame_bytes,
            "frame_hbm_frac": frame_bytes / (total_ms / args.steps * 1e-3) / 1e9 / hbm_peak,
            "roofline": roofline, "roofline_kernels": roofline_all, "cpu_baseline": cpu, "clocks": clocks,
        }
        if other:
            other["value"] = (scene.rays_per_frame() * other["frames_per_step"] + int(refl.item()) / frames_per_step * other["frames_per_step"]) / (other["ms_per_step"] * 1e-3) / 1e6
            other["unit"] = "Mrays/s"
            line[other["mode"]] = other
        line["invalid_series"] = invalid_all   # series whose last frame overflowed a pair arena on some rank (incomplete frames): must be empty
        if timeline:
            line["timeline_us"] = timeline
        if balance:
            line["strip_balance"] = balance
        if check:
            line["gathered_frame_check"] = check
            line["single_gpu_same_frame"] = single
        if config_d:
            line["config_D_single_gpu"] = config_d
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    r.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=["A", "B", "C", "D", "E"], help="default: B on one GPU, D on several")
    ap.add_argument("--no-config-d", action="store_true", help="N = 1: skip the side measurement of config D")
    ap.add_argument("--bounces", type=int, default=-1, help="reflection bounces (-1: what the config names)")
    ap.add_argument("--shard", default="auto", choices=["auto", "frames", "strips"], help="N > 1: strips = one frame in image strips (strong scaling, default), frames = one frame per GPU (weak)")
    ap.add_argument("--light-grid", default="perspective", choices=["perspective", "spherical"],
                    help="perspective: the light path north-star names (default); spherical: the one the committed display() runs")
    ap.add_argument("--shader", default="lambert", choices=["lambert", "spot"], help="spot: Shader::spotlight_shade (display() from its third frame on)")
    ap.add_argument("--slabs", type=int, default=1, help="NUM_SLABS (main.cu.h:18 compiles 1)")
    ap.add_argument("--equal-strips", action="store_true", help="N > 1: keep the equal split instead of balancing the strip boundaries by measurement")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and world != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
