#!/bin/bash
# GPU job 17 (--gpus 2): the NCCL gradient sync on hardware with the final defaults + the default 2-GPU bench line
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t17_dp.log 2>&1; echo "pytest exit $?" >> gpurun_out/t17_dp.log)
tail -4 gpurun_out/t17_dp.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench17_n2.json 2> gpurun_out/bench17_n2.err)
python - <<PY
import json
s=open("gpurun_out/bench17_n2.json").read(); d=json.loads([l for l in s.splitlines() if l.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["config"]["baseline_config3_256_per_gpu"], d["replicas_identical"], d["config"]["sync"])
PY
