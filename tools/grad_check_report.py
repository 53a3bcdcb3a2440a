#!/usr/bin/env python3
"""run_grad_check (src/grad_check.cc:20-61) on reference-shaped nets, as a table: per weighted edge the first epsilon of
[1e-2, 1e-3, 1e-4] that passes (mean |analytic - numeric| / scale < 0.01 over the first 10 weights / biases), or the best one.

    python tools/grad_check_report.py [model+gradcheck ...]      default: lenet (BASELINE config 1 shapes, batch 100), tiny, gradcheck
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convnet_b200 import lib  # noqa: E402
from convnet_b200.net import Net  # noqa: E402

lib.load()
lib.set_precision("fp32")
specs = [(a, 100 if a.startswith("lenet") else 8) for a in sys.argv[1:]] or [("lenet+gradcheck", 100), ("tiny+gradcheck", 32), ("gradcheck", 8)]
print("| net | batch | seed | edge | epsilon | mean scaled diff (weights) | (bias) | pass (< 0.01) |")
print("|---|---|---|---|---|---|---|---|")
for model, batch in specs:
    for seed in (1, 5, 9):
        n = Net(model, batch, seed=3, grad_checker=True)
        for name, eps, dw, db in n.grad_check(seed=seed):
            print("| %s | %d | %d | %s | %g | %.2e | %.2e | %s |" % (model, batch, seed, name, eps, dw, db,
                                                                   "yes" if dw < 0.01 and db < 0.01 else "NO"))
        n.close()
