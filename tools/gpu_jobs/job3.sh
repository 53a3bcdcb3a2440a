#!/bin/bash
# GPU job 3: dgrad in fprop form + fused pool bias-grad: quick probe, full suite, layer probe, bench, launch list
mkdir -p gpurun_out
(MODE=bf16 FAST_REF=1 ITERS=5 timeout 300 python tools/tc_probe.py n128 stride2 conv3 conv5 conv2 > gpurun_out/fast_probe3.log 2>&1; echo "probe exit $?" >> gpurun_out/fast_probe3.log)
tail -20 gpurun_out/fast_probe3.log
(timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/t3.log 2>&1; echo "pytest exit $?" >> gpurun_out/t3.log)
tail -15 gpurun_out/t3.log
(timeout 300 python tools/layer_probe.py > gpurun_out/probe_fast3.log 2>&1)
cat gpurun_out/probe_fast3.log
(timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench3_n1.json 2> gpurun_out/bench3_n1.err)
tail -c 600 gpurun_out/bench3_n1.json
(PRECISION=bf16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches3.csv python tools/step_once.py > gpurun_out/step_once3.log 2>&1)
python tools/launch_summary.py gpurun_out/step_launches3.csv > gpurun_out/step_launches3.md 2>&1; head -40 gpurun_out/step_launches3.md
