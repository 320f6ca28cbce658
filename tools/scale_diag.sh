#!/bin/bash
# development aid: config D in N strips - the bench line, then the same with per-rank timeline markers (UG_TIMELINE=1) and with env variants
# usage: tools/scale_diag.sh N tag [ENV=VAL ...]
N=$1; TAG=$2; shift 2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
O=gpurun_out
env "$@" $TR bench.py --gpus $N --steps 30 --warmup 5 > $O/${TAG}_n$N.json 2> $O/${TAG}_n$N.err
env "$@" UG_TIMELINE=1 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_n${N}_tl.json 2> $O/${TAG}_n${N}_tl.err
python - $O/${TAG}_n$N.json $O/${TAG}_n${N}_tl.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "p50", round(d["ms_per_frame_p50"],4), "single", d.get("single_gpu_same_frame",{}).get("ms_per_frame"), "check", d.get("gathered_frame_check",{}).get("equal"), "frames-mode", (d.get("frames") or {}).get("ms_per_step"))
print(" all", d.get("ms_per_frame_all"))
print(" stage", d["stage_ms"], "e2e", d["e2e"]["ms_per_step"])
try:
    t=json.load(open(sys.argv[2]))
    for k,v in t["timeline_us"].items(): print("  rank",k,v)
except Exception as e: print("no timeline", e)
PY
