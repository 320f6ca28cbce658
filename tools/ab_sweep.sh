#!/bin/bash
# development aid (one GPU): parity tests, then configs with the group-wide look-back of the sort passes (default build) against
# the one-lane walk (lib_nowide.so) and three resident blocks per SM for the ray -> light cell map (lib_rgA3.so)
O=gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > $O/ab_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/ab_pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); st=d.get("stage_ms",{})
    print(f"{sys.argv[1]:26s} {d['ms_per_step']:.4f} ms p50 {d.get('ms_per_frame_p50',0):.4f} e2e {d['e2e']['ms_per_step']:.4f} cam {st.get('cam_grid',0):.4f} lgrid {st.get('light0_grid',0):.4f} regroup {st.get('light0_regroup',0):.4f} shade {st.get('shade',0):.4f} shadow {st.get('light0_shadow',0):.4f} trace {st.get('trace',0):.4f}")
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
}
run() { # name lib args
  UGRT_LIB=$PWD/uniformgrid-raytracing_b200/$2 timeout 300 python bench.py --no-cpu-baseline $3 > $O/ab_$1.json 2> $O/ab_$1.err; show "$1" $O/ab_$1.json
}
for c in B A C D; do
  a="--config $c --steps 40 --warmup 5"; [ $c = B ] && a="--no-config-d"
  run ${c}_wide libugrt.so "$a"
  run ${c}_nowide lib_nowide.so "$a"
  run ${c}_rgA3 lib_rgA3.so "$a"
done
run B_nobounce_wide libugrt.so "--no-config-d --bounces 0"
run B_nobounce_nowide lib_nowide.so "--no-config-d --bounces 0"
