#!/bin/bash
# development aid (one GPU): parity tests, config D / B with and without the per-triangle shading normals, then full ncu
# captures of config D's heaviest kernels (one launch each)
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/ab_pytest_gpu.log 2>&1; tail -3 $O/ab_pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); st=d.get("stage_ms",{})
    print(f"{sys.argv[1]:34s} {d['ms_per_step']:.4f} ms p50 {d.get('ms_per_frame_p50',0):.4f} e2e {d['e2e']['ms_per_step']:.4f} regroup {st.get('light0_regroup',0):.4f} shade {st.get('shade',0):.4f} shadow {st.get('light0_shadow',0):.4f} trace {st.get('trace',0):.4f} cam {st.get('cam_grid',0):.4f}")
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
}
for c in D C; do
UG_NO_HOIST_NORMALS=1 python bench.py --config $c --no-cpu-baseline --steps 30 --warmup 5 > $O/ab_${c}_nohoist.json 2> $O/ab_${c}_nohoist.err; show "$c per-pixel normals" $O/ab_${c}_nohoist.json
python bench.py --config $c --no-cpu-baseline --steps 30 --warmup 5 > $O/ab_${c}_hoist.json 2> $O/ab_${c}_hoist.err; show "$c per-triangle normals" $O/ab_${c}_hoist.json
done
UG_NO_HOIST_NORMALS=1 python bench.py --no-cpu-baseline --no-config-d > $O/ab_B_nohoist.json 2> $O/ab_B_nohoist.err; show "B per-pixel normals" $O/ab_B_nohoist.json
python bench.py --no-cpu-baseline --no-config-d > $O/ab_B_hoist.json 2> $O/ab_B_hoist.err; show "B per-triangle normals" $O/ab_B_hoist.json
for k in k_shadow k_trace k_regroup k_shade; do
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 3 -c 1 -o $O/r02_prof_D_$k -f python bench.py --config D --no-cpu-baseline --steps 3 --warmup 3 > $O/ncu_D_$k.log 2>&1
done
ls -la $O | grep prof_D
