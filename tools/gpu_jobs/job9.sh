#!/bin/bash
# GPU job 9: tests + bench on the current build; ncu captures for profiles; compute-sanitizer on unit shapes
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t9.log 2>&1; echo "pytest exit $?" >> gpurun_out/t9.log)
tail -6 gpurun_out/t9.log
(timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench9_n1.json 2> gpurun_out/bench9_n1.err)
tail -c 1200 gpurun_out/bench9_n1.json
(timeout 300 python tools/layer_probe.py > gpurun_out/probe_fast9.log 2>&1); grep -E "fc|TOTAL|conv1" gpurun_out/probe_fast9.log
(CONVNET_B200_RNORM_TL=64 timeout 120 python tools/membw_probe.py > gpurun_out/membw_tl64.log 2>&1); (timeout 120 python tools/membw_probe.py > gpurun_out/membw_default.log 2>&1); tail -12 gpurun_out/membw_default.log
(PRECISION=bf16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches9.csv python tools/step_once.py > gpurun_out/step_once9.log 2>&1)
python tools/launch_summary.py gpurun_out/step_launches9.csv > gpurun_out/step_launches9.md 2>&1; head -34 gpurun_out/step_launches9.md
# headline kernel, full ncu set (conv4 fprop batch 256: the roofline leg's kernel); -k fast kernel, skip the warm-up launches
(PRECISION=bf16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_fast_kernel -s 6 -c 1 -f -o gpurun_out/r2_conv4_fprop_fast python tools/roofline_probe.py > gpurun_out/ncu_conv4.log 2>&1)
# sanitizer passes on unit shapes (general + lean kernels, pool, rnorm)
(timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "test_conv_vs_oracle and (testconv_full or cin24_s2 or batch_260 or fc_splitk_b32) or test_pool_vs_oracle and ragged or test_rnorm_vs_oracle and testconv" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/sanitizer_memcheck.log)
tail -8 gpurun_out/sanitizer_memcheck.log
(timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "test_conv_vs_oracle and (testconv_full or cin24_s2) and bf16 or test_rnorm_vs_oracle and testconv" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/sanitizer_racecheck.log)
tail -8 gpurun_out/sanitizer_racecheck.log
