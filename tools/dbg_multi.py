import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch.distributed as dist
import ugrt, uniformgrid_raytracing_b200 as ug
from uniformgrid_raytracing_b200 import scenes
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
scene=scenes.scene_d()
r=ug.Renderer(local); r.set_scene(scene); ws = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ws); r.set_stream(ws.cuda_stream)
uid=[ug.unique_id() if rank==0 else None]; dist.broadcast_object_list(uid, src=0); r.comm_init(rank, world, uid[0])
host=torch.empty(3*scene.num_pixels,dtype=torch.uint8).pin_memory()
r.render_host(r.params(scene, strip=(rank,world)), host.numpy())
for i in range(5): r.render(r.params(scene, frame=i, strip=(rank,world)))
torch.cuda.synchronize(); dist.barrier()
for mode in ("nobarrier","barrier"):
    evs=[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    for i in range(8):
        if mode=="barrier":
            torch.cuda.synchronize(); dist.barrier()
        evs[i][0].record()
        r.render(r.params(scene, frame=i, strip=(rank,world)))
        evs[i][1].record()
    torch.cuda.synchronize()
    print(rank, mode, [round(a.elapsed_time(b),3) for a,b in evs], "lib ms_frame", round(r.stats().ms_frame,3), flush=True)
    dist.barrier()
dist.destroy_process_group()
