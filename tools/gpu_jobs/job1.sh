#!/bin/bash
# GPU job 1: new parity tests (reference CUDA kernels, ABI-2, drop-in), layer probes, ncu captures of epilogue-bound kernels
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t1.log 2>&1; echo "pytest exit $?" >> gpurun_out/t1.log)
(timeout 300 python tools/layer_probe.py > gpurun_out/probe_default.log 2>&1)
(CONVNET_B200_2CTA_OPS=7 OPS=dgrad timeout 300 python tools/layer_probe.py > gpurun_out/probe_dgrad_pair.log 2>&1)
for spec in "nin2_1 dgrad" "nin2_1 fprop" "conv2 dgrad"; do
  set -- $spec
  (ONLY=1 OPS=$2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 1 -f -o gpurun_out/$1_$2 python tools/layer_probe.py $1 > gpurun_out/ncu_$1_$2.log 2>&1)
done
tail -5 gpurun_out/t1.log; cat gpurun_out/probe_default.log
