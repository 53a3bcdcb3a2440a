#!/usr/bin/env python3
"""Achieved (algorithmic) GB/s of the HBM-bound kernels at the AlexNet shapes, CUDA-event timed, L2 flushed."""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convnet_b200 import conv_gemm as cg, lib
from convnet_b200.abi import GetConvDesc, num_modules
from convnet_b200.matrix import CUDAMatrix
lib.load()
N = int(os.environ.get("BATCH", "128"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, iters=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
def rnd(r, c, s4):
    m = CUDAMatrix(r, c, s4); m.storage.normal_(); return m
# copy reference
x = torch.empty(N * 110 * 110 * 96, device="cuda"); y = torch.empty_like(x)
t = timed(lambda: y.copy_(x)); print("torch copy %d MB: %.0f GB/s" % (x.numel() * 4 >> 20, 2 * x.numel() * 4 / t / 1e6))
for name, W, C, k, s, p in (("pool1", 110, 96, 3, 2, 1), ("pool2", 27, 256, 3, 2, 1), ("pool5", 12, 512, 3, 2, 1)):
    mod = num_modules(W, k, s, p); d = GetConvDesc(C, C, k, k, s, s, p, p)
    ish, psh = (N, W, W, C), (N, mod, mod, C)
    im, out, gr, tg = rnd(N, W * W * C, ish), CUDAMatrix(N, mod * mod * C, psh), rnd(N, mod * mod * C, psh), CUDAMatrix(N, W * W * C, ish)
    t = timed(lambda: cg.MaxPool(im, out, d)); b = 4 * N * (W * W * C + mod * mod * C)
    print("%s fwd   %7.1f us %6.0f GB/s" % (name, t * 1e3, b / t / 1e6))
    t = timed(lambda: cg.MaxPoolUndo(im, gr, out, tg, d)); b = 4 * N * (2 * W * W * C + 2 * mod * mod * C)
    print("%s undo  %7.1f us %6.0f GB/s" % (name, t * 1e3, b / t / 1e6))
for name, W, C, k in (("rnorm1", 55, 96, 24), ("rnorm2", 14, 256, 64)):
    ish = (N, W, W, C)
    im, out, gr = rnd(N, W * W * C, ish), CUDAMatrix(N, W * W * C, ish), rnd(N, W * W * C, ish)
    t = timed(lambda: cg.ResponseNormCrossMap(im, out, k, 5e-4, 0.75, False)); b = 4 * N * 2 * W * W * C
    print("%s fwd  %7.1f us %6.0f GB/s" % (name, t * 1e3, b / t / 1e6))
    t = timed(lambda: cg.ResponseNormCrossMapUndo(gr, im, out, k, 5e-4, 0.75, False)); b = 4 * N * 3 * W * W * C
    print("%s undo %7.1f us %6.0f GB/s" % (name, t * 1e3, b / t / 1e6))
L = lib.load()
a = torch.randn(N * 110 * 110 * 96, device="cuda")
t = timed(lambda: L.cnb_relu(a.data_ptr(), a.numel())); print("relu (r+w) %7.1f us %6.0f GB/s" % (t * 1e3, 8 * a.numel() / t / 1e6))
w, h, g = (torch.randn(104321024, device="cuda") for _ in range(3))
t = timed(lambda: L.cnb_sgd_momentum(w.data_ptr(), h.data_ptr(), g.data_ptr(), w.numel(), 0.01, 0.9, 5e-4)); print("sgd (3r+2w) %7.1f us %6.0f GB/s" % (t * 1e3, 20 * w.numel() / t / 1e6))
