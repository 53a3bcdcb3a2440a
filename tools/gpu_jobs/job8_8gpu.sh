#!/bin/bash
# GPU job 8 (--gpus 8): hardware DP test at 4 ranks, scaling sweep 1/2/4/8 with the side/comm-stream pipeline, NCCL CTA cap on/off
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t8_dp.log 2>&1; echo "pytest exit $?" >> gpurun_out/t8_dp.log)
tail -4 gpurun_out/t8_dp.log
runN() { n=$1; tag=$2; shift 2; (env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $n --steps 40 --warmup 5 > gpurun_out/bench8_n${n}_$tag.json 2> gpurun_out/bench8_n${n}_$tag.err); python - <<PY
import json
try:
    s=open("gpurun_out/bench8_n${n}_$tag.json").read(); d=json.loads([l for l in s.splitlines() if l.startswith("{")][-1])
    c3=d["config"].get("baseline_config3_256_per_gpu") or {}
    print("n=$n $tag", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "cfg3", round(c3.get("images_per_s",0)), d.get("replicas_identical"))
except Exception as e: print("n=$n $tag failed", e)
PY
}
(timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench8_n1.json 2> gpurun_out/bench8_n1.err); python -c "
import json; d=json.load(open('gpurun_out/bench8_n1.json')); print('n=1', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))"
runN 8 ctas8 CONVNET_B200_NCCL_CTAS=8
runN 8 nocap CONVNET_B200_NCCL_CTAS=0
runN 4 ctas8 CONVNET_B200_NCCL_CTAS=8
runN 2 ctas8 CONVNET_B200_NCCL_CTAS=8
tail -3 gpurun_out/bench8_n8_ctas8.err
