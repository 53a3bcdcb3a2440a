#!/bin/bash
# GPU job 2: validate the lean (fast) conv kernels, then the whole GPU suite, probes, and a step launch list
mkdir -p gpurun_out
(MODE=bf16 FAST_REF=1 ITERS=5 timeout 300 python tools/tc_probe.py n128 conv3 conv4 conv5 conv2 > gpurun_out/fast_probe.log 2>&1; echo "probe exit $?" >> gpurun_out/fast_probe.log)
tail -20 gpurun_out/fast_probe.log
(timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/t2.log 2>&1; echo "pytest exit $?" >> gpurun_out/t2.log)
tail -15 gpurun_out/t2.log
(timeout 300 python tools/layer_probe.py > gpurun_out/probe_fast.log 2>&1)
cat gpurun_out/probe_fast.log
(CONVNET_B200_NO_FAST=1 timeout 300 python tools/layer_probe.py conv2 conv3 conv4 nin2_1 nin4_2 > gpurun_out/probe_nofast.log 2>&1)
(timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err)
tail -c 1500 gpurun_out/bench_n1.json
(PRECISION=bf16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches.csv python tools/step_once.py > gpurun_out/step_once.log 2>&1)
python tools/launch_summary.py gpurun_out/step_launches.csv > gpurun_out/step_launches.md 2>&1; head -40 gpurun_out/step_launches.md
