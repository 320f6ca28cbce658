#!/bin/bash
# development aid: programmatic dependent launch A/B on one GPU.  Builds compared (all in-tree, built by csrc/build.sh):
#   libugrt.so       the default build
#   libugrt_pdl.so   UG_EXTRA_FLAGS=-DUG_PDL=1 (griddepcontrol.wait + launch_dependents at the head of the chain's kernels)
#   libugrt_pdl2.so  UG_EXTRA_FLAGS=-DUG_PDL=2 (wait only: the dependents go when the last block has exited)
# Bench lines land in gpurun_out/pdl_*.json; if the PDL build cannot capture its frame graph the kernel groups are tried one
# by one (UG_PDL_MASK), else the GPU parity tests run on the PDL build.
O=gpurun_out
P=$PWD/uniformgrid-raytracing_b200
T0=$(date +%s)
run() { # name lib "bench args" [env...]
  local name=$1 lib=$2 args=$3; shift 3
  env "$@" UGRT_LIB=$P/$lib timeout 90 python bench.py --no-cpu-baseline --no-config-d $args > $O/pdl_$name.json 2> $O/pdl_$name.err
  local rc=$?
  python - "$name" $O/pdl_$name.json $rc $(( $(date +%s) - T0 )) <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); e=d.get("e2e",{})
    print(f"{sys.argv[1]:18s} rc={sys.argv[3]} t+{sys.argv[4]}s {d['ms_per_step']:.4f} ms p50 {d.get('ms_per_frame_p50',0):.4f} p10 {d.get('ms_per_frame_p10',0):.4f} e2e {e.get('ms_per_step',0):.4f} sync {e.get('one_frame_at_a_time',{}).get('ms_per_step',0):.4f} {d.get('library',{}).get('version')} invalid={d.get('invalid_series')}")
except Exception as ex: print(sys.argv[1], "rc=" + sys.argv[3], "ERR", ex)
PY
  return $rc
}
run B_plain libugrt.so ""
if run B_pdl libugrt_pdl.so ""; then
  run B_pdl2 libugrt_pdl2.so ""
  run B_plain_b libugrt.so ""
  run B_pdl_b libugrt_pdl.so ""
  UGRT_LIB=$P/libugrt_pdl.so timeout 200 python -m pytest tests -x -q -m gpu > $O/pdl_pytest_gpu.log 2>&1; tail -3 $O/pdl_pytest_gpu.log; echo "t+$(( $(date +%s) - T0 ))s"
  run D_plain libugrt.so "--config D --steps 30 --warmup 5"
  run D_pdl libugrt_pdl.so "--config D --steps 30 --warmup 5"
  run B0_plain libugrt.so "--bounces 0"
  run B0_pdl libugrt_pdl.so "--bounces 0"
  run A_plain libugrt.so "--config A"
  run A_pdl libugrt_pdl.so "--config A"
else
  tail -5 $O/pdl_B_pdl.err
  U=0
  for m in 1 2 4 8 16 32 64; do
    if run B_mask$m libugrt_pdl.so "--steps 30 --warmup 5" UG_PDL_MASK=$m; then U=$((U | m)); else tail -2 $O/pdl_B_mask$m.err; fi
  done
  echo "groups that capture: mask $U"
  if [ $U -ne 0 ]; then
    run B_union libugrt_pdl.so "" UG_PDL_MASK=$U
    run B_plain_b libugrt.so ""
    UG_PDL_MASK=$U UGRT_LIB=$P/libugrt_pdl.so timeout 200 python -m pytest tests -x -q -m gpu > $O/pdl_pytest_gpu.log 2>&1; tail -3 $O/pdl_pytest_gpu.log
  fi
fi
echo "total $(( $(date +%s) - T0 ))s"
