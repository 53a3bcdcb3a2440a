#!/bin/bash
# GPU job 18: trimmed pool-undo patch kernel — pool / staging / net tests, then the bench line
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_staging.py tests/test_gpu_parity.py tests/test_gpu_net.py -m gpu -q --tb=short -p no:cacheprovider -k "pool or Pool or staging or alexnet or traced" > gpurun_out/t18.log 2>&1; echo "pytest exit $?" >> gpurun_out/t18.log)
tail -4 gpurun_out/t18.log
(timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench18_n1.json 2> gpurun_out/bench18_n1.err)
python - <<PY
import json
s=open("gpurun_out/bench18_n1.json").read(); d=json.loads([l for l in s.splitlines() if l.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["roofline"]["frac"], d["clocks"])
PY
