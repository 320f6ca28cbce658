#!/bin/bash
# development aid (one GPU): parity tests, then config D and B with the stored / recomputed primary ray direction
# (UG_DIR_RECOMPUTE: bit 0 k_shade, bit 1 ray -> light cell map) and a build with 4 pixels per trip in the map
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/ab_pytest_gpu.log 2>&1; tail -3 $O/ab_pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); st=d.get("stage_ms",{})
    print(f"{sys.argv[1]:34s} {d['ms_per_step']:.4f} ms p50 {d.get('ms_per_frame_p50',0):.4f} e2e {d['e2e']['ms_per_step']:.4f} regroup {st.get('light0_regroup',0):.4f} shade {st.get('shade',0):.4f} shadow {st.get('light0_shadow',0):.4f} trace {st.get('trace',0):.4f}")
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
}
for m in 0 1 2 3; do
  UG_DIR_RECOMPUTE=$m python bench.py --config D --no-cpu-baseline --steps 30 --warmup 5 > $O/ab_D_dir$m.json 2> $O/ab_D_dir$m.err; show "D UG_DIR_RECOMPUTE=$m" $O/ab_D_dir$m.json
done
UGRT_LIB=$PWD/uniformgrid-raytracing_b200/lib_ka4.so python bench.py --config D --no-cpu-baseline --steps 30 --warmup 5 > $O/ab_D_ka4.json 2> $O/ab_D_ka4.err; show "D RG_KA=4 (recompute 3)" $O/ab_D_ka4.json
for m in 0 3; do
  UG_DIR_RECOMPUTE=$m python bench.py --no-cpu-baseline --no-config-d > $O/ab_B_dir$m.json 2> $O/ab_B_dir$m.err; show "B UG_DIR_RECOMPUTE=$m" $O/ab_B_dir$m.json
done
UGRT_LIB=$PWD/uniformgrid-raytracing_b200/lib_ka4.so python bench.py --no-cpu-baseline --no-config-d > $O/ab_B_ka4.json 2> $O/ab_B_ka4.err; show "B RG_KA=4 (recompute 3)" $O/ab_B_ka4.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/ab_B_dir3.json"))
print("fp32_issue", d["roofline_kernels"].get("k_trace",{}).get("fp32_issue"))
PY
