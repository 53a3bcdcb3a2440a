#!/bin/bash
# GPU job 6: tests with side-lane bias grads + tiled banks + NVTX; bench; other BASELINE configs; grad-check report
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t6.log 2>&1; echo "pytest exit $?" >> gpurun_out/t6.log)
tail -6 gpurun_out/t6.log
(timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench6_n1.json 2> gpurun_out/bench6_n1.err)
tail -c 300 gpurun_out/bench6_n1.json
(CONVNET_B200_NO_EAGER_UPDATE=1 CONVNET_B200_NO_SIDE_BIAS_GRAD=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench6_n1_noside.json 2>/dev/null)
tail -c 300 gpurun_out/bench6_n1_noside.json
(timeout 300 python tools/bench_configs.py train --model alexnet --batch 256 > gpurun_out/cfg3_b256.json 2> gpurun_out/cfg3.err); cat gpurun_out/cfg3_b256.json
(timeout 300 python tools/bench_configs.py train --model c3d --batch 32 > gpurun_out/cfg4_c3d.json 2> gpurun_out/cfg4.err); cat gpurun_out/cfg4_c3d.json; tail -3 gpurun_out/cfg4.err
(timeout 300 python tools/bench_configs.py infer --model alexnet --batch 512 > gpurun_out/cfg5_infer.json 2> gpurun_out/cfg5.err); cat gpurun_out/cfg5_infer.json
(timeout 600 python tools/grad_check_report.py > gpurun_out/grad_check_report.md 2> gpurun_out/grad_check.err); cat gpurun_out/grad_check_report.md | head -40
