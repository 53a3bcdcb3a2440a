#!/usr/bin/env python3
"""Secondary BASELINE.json configurations, one JSON line each (not the driver's bench line — that is bench.py):

    python tools/bench_configs.py train --model alexnet --batch 256      # cfg3 per-GPU shape (batch 256 / GPU)
    python tools/bench_configs.py train --model c3d --batch 32           # cfg4: C3D-style clips, 16 x 112 x 112 x 3
    python tools/bench_configs.py infer --model alexnet --batch 512      # cfg5: extract_representation, fprop only

Same timing rules as bench.py: >= 3 warm-up steps, CUDA events on the launching stream, inputs resident in HBM.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convnet_b200 import lib  # noqa: E402
from convnet_b200.net import Net  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["train", "infer"])
    ap.add_argument("--model", default="alexnet")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "tf32", "bf16"])
    args = ap.parse_args()
    lib.load(); lib.set_precision(args.precision)
    n = Net(args.model, args.batch, seed=1)
    g = torch.Generator(device="cuda").manual_seed(1234)
    n.input_tensor().normal_(generator=g)
    n.labels_tensor().copy_(torch.randint(0, n.num_classes, (args.batch,), device="cuda", generator=g, dtype=torch.int32))
    step = (lambda: n.train_step(False)) if args.mode == "train" else (lambda: n.fprop(False))
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    flops = n.flops_train if args.mode == "train" else n.flops_fprop
    print(json.dumps({"mode": args.mode, "model": args.model, "batch": args.batch, "precision": args.precision,
                      "ms_per_step": ms, "items_per_s": args.batch / (ms * 1e-3), "model_tflops": flops / (ms * 1e-3) / 1e12,
                      "gflop_per_item": flops / args.batch / 1e9, "steps": args.steps}))
    n.close()


if __name__ == "__main__":
    main()
