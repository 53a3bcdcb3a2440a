#!/bin/bash
# GPU job 10 (--gpus 2): pool patch kernels + 3-stream pipeline; step timelines at N=1 and N=2; NCCL all-reduce alone
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_staging.py tests/test_gpu_parity.py tests/test_gpu_net.py -m gpu -q --tb=short -p no:cacheprovider -k "pool or Pool or test_gpu_net or staging" > gpurun_out/t10.log 2>&1; echo "pytest exit $?" >> gpurun_out/t10.log)
tail -5 gpurun_out/t10.log
show() { python - <<PY
import json
try:
    s=open("gpurun_out/$1.json").read(); d=json.loads([l for l in s.splitlines() if l.startswith("{")][-1])
    print("$1", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), (d["config"].get("baseline_config3_256_per_gpu") or {}).get("images_per_s"))
    print(json.dumps(d.get("timeline_rank0")))
except Exception as e: print("$1 failed", e)
PY
}
(timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench10_n1.json 2> gpurun_out/bench10_n1.err); show bench10_n1
(CONVNET_B200_POOL_PATCH=0 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench10_n1_nopatch.json 2> gpurun_out/bench10_n1_nopatch.err); show bench10_n1_nopatch
run2() { tag=$1; shift; (env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench10_n2_$tag.json 2> gpurun_out/bench10_n2_$tag.err); show bench10_n2_$tag; }
run2 ctas8 CONVNET_B200_NCCL_CTAS=8
run2 ctas32 CONVNET_B200_NCCL_CTAS=32
for c in 0 8 32; do
  (if [ $c -gt 0 ]; then export NCCL_MAX_CTAS=$c; fi; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/allreduce_probe.py 2>/dev/null | grep busbw | sed "s/^/ctas$c /" | tee -a gpurun_out/allreduce_n2.log)
done
(timeout 120 python tools/membw_probe.py > gpurun_out/membw10.log 2>&1); grep pool gpurun_out/membw10.log
