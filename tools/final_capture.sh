#!/bin/bash
# One-GPU validation + evidence capture for a round: parity tests, smoke, bench lines per config, ncu launch lists and
# full captures of the traversal kernels and one sort pass.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
R=${1:-r02}
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > $O/${R}_pytest_gpu.log 2>&1; tail -2 $O/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${R}_smoke.log 2>&1; tail -1 $O/${R}_smoke.log
python bench.py > $O/${R}_bench_B.json 2> $O/${R}_bench_B.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_B_reference.json 2> $O/${R}_bench_B_reference.err
python bench.py --bounces 0 --no-cpu-baseline --no-config-d > $O/${R}_bench_B_nobounce.json 2>/dev/null
python bench.py --config A --no-cpu-baseline > $O/${R}_bench_A.json 2>/dev/null
python bench.py --config C --no-cpu-baseline --steps 50 > $O/${R}_bench_C.json 2>/dev/null
python bench.py --config D --no-cpu-baseline --steps 30 --warmup 5 > $O/${R}_bench_D.json 2>/dev/null
python bench.py --light-grid spherical --shader spot --bounces 0 --no-cpu-baseline --no-config-d --steps 50 > $O/${R}_bench_B_display_faithful.json 2>/dev/null
[ -n "$WITH_E" ] && python bench.py --config E --no-cpu-baseline --steps 5 --warmup 3 > $O/${R}_bench_E.json 2>/dev/null
for f in B B_reference B_nobounce B_display_faithful A C D E; do [ -f $O/${R}_bench_$f.json ] && python - $O/${R}_bench_$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), d.get("stage_ms"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
# profiler passes only after the plain runs above exited 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_B.csv python bench.py --no-cpu-baseline --no-config-d --steps 3 --warmup 3 > $O/ncu_l.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${R}_launches_D.csv python bench.py --config D --no-cpu-baseline --steps 3 --warmup 3 > $O/ncu_l.log 2>&1
python tools/launch_summary.py $O/${R}_launches_B.csv > $O/${R}_launches_B_summary.txt; cat $O/${R}_launches_B_summary.txt
python tools/launch_summary.py $O/${R}_launches_D.csv > $O/${R}_launches_D_summary.txt
for k in ${FULL_KERNELS:-k_trace k_refl_trace k_shadow k_sweep_pass k_sweep_setup k_regroup k_shade}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $O/${R}_prof_$k -f python bench.py --no-cpu-baseline --no-config-d --steps 3 --warmup 3 > $O/ncu_$k.log 2>&1
done
ls -la $O | grep ${R}_ | head -60
