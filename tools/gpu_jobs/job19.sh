#!/bin/bash
# GPU job 19: the input-pipeline kernel and host iterator against the oracle
mkdir -p gpurun_out
(timeout 100 python -m pytest tests/test_input_pipeline.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t19.log 2>&1; echo "pytest exit $?" >> gpurun_out/t19.log)
tail -25 gpurun_out/t19.log
