#!/bin/bash
# development aid: programmatic dependent launch A/B on one GPU (k_common.cuh).  Builds compared, all in-tree:
#   libugrt.so        the default build
#   libugrt_e<M>.so   UG_EXTRA_FLAGS="-DUG_PDL=1 -DUG_PDL_EARLY=<M>" csrc/build.sh: every kernel of the chain waits first
#                     (griddepcontrol.wait); the groups in mask M also let their dependents go early (launch_dependents)
# usage: tools/ab_pdl.sh "<masks>"   e.g. "0 1 2 4 8 16 32 64 5 85 87 119"
# Config B decides (lowest ms/frame); the GPU parity tests, the default bench line and the other configs then run on that build.
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
    print(f"{sys.argv[1]:14s} rc={sys.argv[3]} t+{sys.argv[4]}s {d['ms_per_step']:.4f} ms p50 {d.get('ms_per_frame_p50',0):.4f} p10 {d.get('ms_per_frame_p10',0):.4f} e2e {e.get('ms_per_step',0):.4f} sync {e.get('one_frame_at_a_time',{}).get('ms_per_step',0):.4f} invalid={d.get('invalid_series')}")
except Exception as ex: print(sys.argv[1], "rc=" + sys.argv[3], "ERR", ex)
PY
  return $rc
}
run B_plain libugrt.so ""
for m in $1; do run B_e$m libugrt_e$m.so ""; done
BEST=$(python - $1 <<'PY'
import json,sys
best=None
for m in sys.argv[1:]:
    try:
        d=json.load(open(f"gpurun_out/pdl_B_e{m}.json"))
        if d.get("invalid_series"): continue
        if best is None or d["ms_per_step"] < best[0]: best=(d["ms_per_step"], m)
    except Exception: pass
print(best[1] if best else "")
PY
)
echo "best early mask on config B: '$BEST'"
[ -z "$BEST" ] && exit 1
L=libugrt_e$BEST.so
UGRT_LIB=$P/$L timeout 200 python -m pytest tests -x -q -m gpu > $O/pdl_pytest_gpu.log 2>&1; tail -2 $O/pdl_pytest_gpu.log; echo "t+$(( $(date +%s) - T0 ))s"
UGRT_LIB=$P/$L python -c "import __graft_entry__ as g; g.smoke()" > $O/pdl_smoke.log 2>&1; tail -1 $O/pdl_smoke.log
UGRT_LIB=$P/$L python bench.py > $O/pdl_bench_B_full.json 2> $O/pdl_bench_B_full.err; echo "full line rc=$? t+$(( $(date +%s) - T0 ))s"
run A_plain libugrt.so "--config A"
run A_best $L "--config A"
run B0_plain libugrt.so "--bounces 0"
run B0_best $L "--bounces 0"
run C_plain libugrt.so "--config C --steps 50"
run C_best $L "--config C --steps 50"
run D_best $L "--config D --steps 30 --warmup 5"
run D_plain libugrt.so "--config D --steps 30 --warmup 5"
echo "total $(( $(date +%s) - T0 ))s"
