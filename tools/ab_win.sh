#!/bin/bash
# development aid (one GPU): parity tests, then the look-back window of the sort passes (one-lane walk: SW_WIN 8 / 16; lanes sharing a digit: SW_WING 4 / 8)
O=gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > $O/ab_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/ab_pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); st=d.get("stage_ms",{})
    print(f"{sys.argv[1]:22s} {d['ms_per_step']:.4f} ms p50 {d.get('ms_per_frame_p50',0):.4f} e2e {d['e2e']['ms_per_step']:.4f} cam {st.get('cam_grid',0):.4f} lgrid {st.get('light0_grid',0):.4f} regroup {st.get('light0_regroup',0):.4f} shadow {st.get('light0_shadow',0):.4f} trace {st.get('trace',0):.4f}")
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
}
run() { # name lib args
  UGRT_LIB=$PWD/uniformgrid-raytracing_b200/$2 timeout 300 python bench.py --no-cpu-baseline $3 > $O/ab_$1.json 2> $O/ab_$1.err; show "$1" $O/ab_$1.json
}
for c in D B C; do
  a="--config $c --steps 40 --warmup 5"; [ $c = B ] && a="--no-config-d"
  run ${c}_base libugrt.so "$a"
  run ${c}_win16 lib_win16.so "$a"
  run ${c}_wing8 lib_wing8.so "$a"
  run ${c}_w16g8 lib_w16g8.so "$a"
done
