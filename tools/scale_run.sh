#!/bin/bash
# development aid: the default bench line (config B, frames mode + strips beside it) and config D in strips at N GPUs
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/r01_scale_B_n$N.json 2> gpurun_out/r01_scale_B_n$N.err
$TR bench.py --gpus $N --config D --shard strips --steps 30 --warmup 5 > gpurun_out/r01_scale_D_strips_n$N.json 2> gpurun_out/r01_scale_D_strips_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
for f in (f"gpurun_out/r01_scale_B_n{n}.json", f"gpurun_out/r01_scale_D_strips_n{n}.json"):
    try:
        d=json.load(open(f)); print(f, d["n_gpus"], d["scaling"], round(d["ms_per_step"],4), round(d["value"]), d.get("strips"), d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
