#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t5.log 2>&1; echo "pytest exit $?" >> gpurun_out/t5.log)
tail -8 gpurun_out/t5.log
(timeout 300 python tools/layer_probe.py > gpurun_out/probe_fast5.log 2>&1)
cat gpurun_out/probe_fast5.log
(timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench5_n1.json 2> gpurun_out/bench5_n1.err)
tail -c 300 gpurun_out/bench5_n1.json
(PRECISION=bf16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches5.csv python tools/step_once.py > gpurun_out/step_once5.log 2>&1)
python tools/launch_summary.py gpurun_out/step_launches5.csv > gpurun_out/step_launches5.md 2>&1; head -36 gpurun_out/step_launches5.md
