#!/bin/bash
# development aid: config E (1 M-triangle soup, 8192x8192, 4 lights, 2 bounces) in N strips; usage: tools/scale_E.sh N [steps]
N=${1:-8}; K=${2:-10}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 420 $TR bench.py --gpus $N --config E --steps $K --warmup 3 > gpurun_out/r02_scale_E_strips_n$N.json 2> gpurun_out/r02_scale_E_strips_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/r02_scale_E_strips_n{n}.json"))
    print("E n",n,"ms/step", d["ms_per_step"], "value", d["value"], "single", d.get("single_gpu_same_frame"), "check", d.get("gathered_frame_check"), "frames", (d.get("frames") or {}).get("ms_per_step"), "invalid", d.get("invalid_series"))
    print(d["stage_ms"]); print(d["config"]["pairs_camera_grid"], d["config"]["pairs_light_grids"])
except Exception as e:
    print("ERR", e); print(open(f"gpurun_out/r02_scale_E_strips_n{n}.err").read()[-3000:])
PY
