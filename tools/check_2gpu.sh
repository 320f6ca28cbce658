#!/bin/bash
# two GPUs: the multi-GPU parity tests, then the driver's own N = 2 command (config D in two image strips, gathered frame hash-checked)
O=gpurun_out
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q > $O/r02_pytest_multi_gpu_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_multi_gpu_2gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 > $O/r02_scale_D_strips_n2.json 2> $O/r02_scale_D_strips_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02_scale_D_strips_n2.json"))
    print(d["n_gpus"], d["scaling"], round(d["ms_per_step"],4), round(d["value"]), d["e2e"]["ms_per_step"], d.get("gathered_frame_check"), d.get("single_gpu_same_frame"), d.get("invalid_series"), d.get("stage_ms"))
    print(d.get("frames"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_scale_D_strips_n2.err").read()[-3000:])
PY
