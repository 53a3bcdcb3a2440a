#!/bin/bash
# GPU job 14 (--gpus 8): scaling on the final build — NCCL width sweep at N=8, then N=4, N=2, N=1 on the same box
mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    s=open("gpurun_out/$1.json").read(); d=json.loads([l for l in s.splitlines() if l.startswith("{")][-1])
    t=d.get("timeline_rank0") or {}
    print("$1", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "cfg3", round(((d["config"].get("baseline_config3_256_per_gpu") or {}).get("images_per_s") or 0)), d.get("replicas_identical"),
          "| trace fwd %.3f bwd %.3f end %.3f |" % (t.get("fprop_end_ms",0), t.get("bprop_compute_end_ms",0), t.get("step_end_ms",0)),
          " ".join("%.0fMB:%.2f-%.2f" % (b["MB"], b["exchange_begin_ms"], b["exchange_end_ms"]) for b in t.get("buckets",[])))
except Exception as e: print("$1 failed", e)
PY
}
runN() { n=$1; tag=$2; mb=$3; extra=$4; shift 4; (env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 30 --warmup 5 --bucket-mb $mb $extra > gpurun_out/bench14_$tag.json 2> gpurun_out/bench14_$tag.err); show bench14_$tag; }
(timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench14_n1.json 2> gpurun_out/bench14_n1.err); show bench14_n1
runN 8 n8_k16_b128 128 "" CONVNET_B200_NCCL_CTAS=16
runN 8 n8_k24_b128 128 --no-cfg3 CONVNET_B200_NCCL_CTAS=24
runN 8 n8_k16_b8 8 --no-cfg3 CONVNET_B200_NCCL_CTAS=16
runN 8 n8_k32_b128 128 --no-cfg3 CONVNET_B200_NCCL_CTAS=32
runN 4 n4_k16_b128 128 --no-cfg3 CONVNET_B200_NCCL_CTAS=16
runN 2 n2_k16_b128 128 --no-cfg3 CONVNET_B200_NCCL_CTAS=16
grep -i "NVLS\|nvls" gpurun_out/bench14_n8_k16_b128.err | head -3
