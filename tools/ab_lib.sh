#!/bin/bash
# development aid: bench one config with several builds of the library; usage: tools/ab_lib.sh "<bench args>" lib1.so lib2.so ...
args="$1"; shift
i=0
for l in "$@"; do
  i=$((i+1))
  UGRT_LIB=$PWD/$l python bench.py --no-cpu-baseline $args > gpurun_out/abl_$i.json 2> gpurun_out/abl_$i.err
  python - "$l" gpurun_out/abl_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); st=d.get("stage_ms",{})
    print(f"{sys.argv[1]:50s} {d['ms_per_step']:.4f} ms p50 {d.get('ms_per_frame_p50',0):.4f} shadow {st.get('light0_shadow',0):.4f} trace {st.get('trace',0):.4f} refl {st.get('reflect_rays',0):.4f} regroup {st.get('light0_regroup',0):.4f}")
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
done
