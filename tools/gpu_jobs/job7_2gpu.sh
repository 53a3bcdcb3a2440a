#!/bin/bash
# GPU job 7 (--gpus 2): hardware test of the NCCL gradient sync + 2-GPU bench with different NCCL CTA caps
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t7_dp.log 2>&1; echo "pytest exit $?" >> gpurun_out/t7_dp.log)
tail -8 gpurun_out/t7_dp.log
run2() { tag=$1; shift; (env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench7_n2_$tag.json 2> gpurun_out/bench7_n2_$tag.err); python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench7_n2_$tag.json")); print("$tag", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["config"].get("baseline_config3_256_per_gpu"))
except Exception as e: print("$tag failed", e)
PY
}
(timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench7_n1.json 2> gpurun_out/bench7_n1.err); python -c "
import json; d=json.load(open('gpurun_out/bench7_n1.json')); print('n1', round(d['value']), d['ms_per_step'])"
run2 ctas8 CONVNET_B200_NCCL_CTAS=8
run2 ctas4 CONVNET_B200_NCCL_CTAS=4
run2 ctas16 CONVNET_B200_NCCL_CTAS=16
run2 nocap CONVNET_B200_NCCL_CTAS=0
run2 noeager CONVNET_B200_NO_EAGER_UPDATE=1
tail -3 gpurun_out/bench7_n2_ctas8.err
