#!/bin/bash
# GPU job 16: final single-GPU validation and artefacts: tests, bench (with cpu_baseline), launch list, ncu --set full of the
# pool / response-norm kernels of one step, memory-bound kernel probe
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t16.log 2>&1; echo "pytest exit $?" >> gpurun_out/t16.log)
tail -6 gpurun_out/t16.log
(timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench16_n1.json 2> gpurun_out/bench16_n1.err)
python - <<PY
import json
s=open("gpurun_out/bench16_n1.json").read(); d=json.loads([l for l in s.splitlines() if l.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["roofline"]["frac"], d["gpu_launches"], d["clocks"], d["cpu_baseline"]["value"])
PY
(PRECISION=bf16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches16.csv python tools/step_once.py > gpurun_out/step_once16.log 2>&1)
python tools/launch_summary.py gpurun_out/step_launches16.csv > gpurun_out/step_launches16.md 2>&1; head -12 gpurun_out/step_launches16.md
(PRECISION=bf16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_undo|pool_fwd|rnorm" --profile-from-start off -f -o gpurun_out/r2_membound python tools/step_once.py > gpurun_out/ncu_membound16.log 2>&1); tail -2 gpurun_out/ncu_membound16.log
(CONVNET_B200_POOL_PATCH=0 PRECISION=bf16 timeout 600 ncu --set full --clock-control none -k regex:"pool_undo" --profile-from-start off -c 1 -s 2 -f -o gpurun_out/r2_pool1_undo_per_element python tools/step_once.py > gpurun_out/ncu_pool_old16.log 2>&1); tail -2 gpurun_out/ncu_pool_old16.log
(timeout 120 python tools/membw_probe.py > gpurun_out/membw16.log 2>&1); cat gpurun_out/membw16.log
