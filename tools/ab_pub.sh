#!/bin/bash
# development aid (one GPU): parity tests, then the counters' publish beside / behind the shading, and the uniform-grid resolution of row R
O=gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > $O/ab_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/ab_pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); st=d.get("stage_ms",{})
    print(f"{sys.argv[1]:30s} {d['ms_per_step']:.4f} ms p50 {d.get('ms_per_frame_p50',0):.4f} e2e {d['e2e']['ms_per_step']:.4f} refl_grid {st.get('reflect_grid',0):.4f} refl_rays {st.get('reflect_rays',0):.4f}")
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
}
run() { # name env args
  env $2 timeout 300 python bench.py --no-cpu-baseline $3 > $O/ab_$1.json 2> $O/ab_$1.err; show "$1" $O/ab_$1.json
}
run B_early X=1 "--no-config-d"
run B_late UG_LATE_PUBLISH=1 "--no-config-d"
run B_early2 X=1 "--no-config-d"
run B_late2 UG_LATE_PUBLISH=1 "--no-config-d"
run A_early X=1 "--config A"
run A_late UG_LATE_PUBLISH=1 "--config A"
run C_early X=1 "--config C --steps 40 --warmup 5"
run C_late UG_LATE_PUBLISH=1 "--config C --steps 40 --warmup 5"
for c in 1.0 1.5 3.0; do run B_cpt$c UG_REFL_CELLS_PER_TRI=$c "--no-config-d"; done
