#!/bin/bash
# GPU job 15: programmatic dependent launch on the lean kernels — tests + A/B bench
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/t15.log 2>&1; echo "pytest exit $?" >> gpurun_out/t15.log)
tail -6 gpurun_out/t15.log
show() { python - <<PY
import json
try:
    s=open("gpurun_out/$1.json").read(); d=json.loads([l for l in s.splitlines() if l.startswith("{")][-1])
    print("$1", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["roofline"]["frac"], d["gpu_launches"])
except Exception as e: print("$1 failed", e)
PY
}
(timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench15_pdl.json 2> gpurun_out/bench15_pdl.err); show bench15_pdl
(CONVNET_B200_NO_PDL=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench15_nopdl.json 2> gpurun_out/bench15_nopdl.err); show bench15_nopdl
(timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench15_pdl2.json 2> gpurun_out/bench15_pdl2.err); show bench15_pdl2
(timeout 300 python tools/layer_probe.py > gpurun_out/probe15.log 2>&1); grep -E "TOTAL" gpurun_out/probe15.log
