#!/bin/bash
# development aid: bench one config under several environment settings; usage: tools/ab_env.sh "<bench args>" "ENV1=.. ENV2=.." "ENV3=.." ...
args="$1"; shift
i=0
for e in "$@"; do
  i=$((i+1))
  env $e python bench.py --no-cpu-baseline $args > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  python - "$e" gpurun_out/ab_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(f"{sys.argv[1]:60s} {d['ms_per_step']:.4f} ms  warm {d.get('ms_per_frame_warm_l2',0):.4f}")
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
done
