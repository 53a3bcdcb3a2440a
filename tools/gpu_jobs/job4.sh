#!/bin/bash
# GPU job 4: x-mode lean kernels, split selection, full suite (no -x), probes, bench, launch list
mkdir -p gpurun_out
(MODE=tf32 FAST_REF=1 ITERS=5 timeout 300 python tools/tc_probe.py conv1_small conv1 mnist1 > gpurun_out/fast_probe4.log 2>&1; echo "probe exit $?" >> gpurun_out/fast_probe4.log)
tail -12 gpurun_out/fast_probe4.log
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t4.log 2>&1; echo "pytest exit $?" >> gpurun_out/t4.log)
tail -15 gpurun_out/t4.log
(timeout 300 python tools/layer_probe.py > gpurun_out/probe_fast4.log 2>&1)
cat gpurun_out/probe_fast4.log
(timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench4_n1.json 2> gpurun_out/bench4_n1.err)
tail -c 400 gpurun_out/bench4_n1.json
(PRECISION=bf16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches4.csv python tools/step_once.py > gpurun_out/step_once4.log 2>&1)
python tools/launch_summary.py gpurun_out/step_launches4.csv > gpurun_out/step_launches4.md 2>&1; head -40 gpurun_out/step_launches4.md
