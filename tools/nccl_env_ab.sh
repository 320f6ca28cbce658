#!/bin/bash
N=$1; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
i=0
for e in "$@"; do
  i=$((i+1))
  env $e $TR bench.py --gpus $N --steps 40 --warmup 10 > gpurun_out/nccl_$i.json 2> gpurun_out/nccl_$i.err
  python - "$e" gpurun_out/nccl_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(f"{sys.argv[1]:70s} frames {d['ms_per_step']:.4f} ms  strips {d['strips']['ms_per_frame']:.4f}")
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
done
