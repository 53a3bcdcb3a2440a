#!/bin/bash
# GPU job 12: full GPU test suite + bench + launch list on the build with pool patches, fused dropout, prestaged banks
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t12.log 2>&1; echo "pytest exit $?" >> gpurun_out/t12.log)
tail -6 gpurun_out/t12.log
show() { python - <<PY
import json
try:
    s=open("gpurun_out/$1.json").read(); d=json.loads([l for l in s.splitlines() if l.startswith("{")][-1])
    print("$1", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["roofline"]["frac"], d["gpu_launches"])
    print(json.dumps(d.get("timeline_rank0")))
except Exception as e: print("$1 failed", e)
PY
}
(timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench12_n1.json 2> gpurun_out/bench12_n1.err); show bench12_n1
(CONVNET_B200_NO_FUSED_DROPOUT=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench12_nodrop.json 2> gpurun_out/bench12_nodrop.err); show bench12_nodrop
(CONVNET_B200_NO_PRESTAGE=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench12_noprestage.json 2> gpurun_out/bench12_noprestage.err); show bench12_noprestage
(PRECISION=bf16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches12.csv python tools/step_once.py > gpurun_out/step_once12.log 2>&1)
python tools/launch_summary.py gpurun_out/step_launches12.csv > gpurun_out/step_launches12.md 2>&1; head -34 gpurun_out/step_launches12.md
