#!/bin/bash
# GPU job 11 (--gpus 2): NCCL width (per-communicator config) x bucket size sweep at N=2, with step timelines
mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    s=open("gpurun_out/$1.json").read(); d=json.loads([l for l in s.splitlines() if l.startswith("{")][-1])
    t=d.get("timeline_rank0") or {}
    print("$1", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "cfg3", round(((d["config"].get("baseline_config3_256_per_gpu") or {}).get("images_per_s") or 0)),
          "| trace fwd %.3f bwd %.3f end %.3f |" % (t.get("fprop_end_ms",0), t.get("bprop_compute_end_ms",0), t.get("step_end_ms",0)),
          " ".join("%.0fMB:%.2f-%.2f" % (b["MB"], b["exchange_begin_ms"], b["exchange_end_ms"]) for b in t.get("buckets",[])))
except Exception as e: print("$1 failed", e)
PY
}
run2() { tag=$1; mb=$2; shift 2; (env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 --bucket-mb $mb --no-cfg3 > gpurun_out/bench11_$tag.json 2> gpurun_out/bench11_$tag.err); show bench11_$tag; }
run2 k16_b8 8 CONVNET_B200_NCCL_CTAS=16
run2 k16_b128 128 CONVNET_B200_NCCL_CTAS=16
run2 k8_b128 128 CONVNET_B200_NCCL_CTAS=8
run2 k24_b128 128 CONVNET_B200_NCCL_CTAS=24
run2 k12_b32 32 CONVNET_B200_NCCL_CTAS=12
run2 k0_b128 128 CONVNET_B200_NCCL_CTAS=0
tail -2 gpurun_out/bench11_k16_b8.err
