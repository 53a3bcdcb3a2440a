#!/usr/bin/env python3
"""Per-layer, per-op device times of the AlexNet (net A) conv-type edges the way the TRAINING STEP calls them:
bf16 operands staged beforehand (convnet_b200_bf16_stage), bias+ReLU fused into fprop, the ReLU' mask fused into dgrad.

    python tools/layer_probe.py [layer ...]            BATCH=128 MODE=bf16 ITERS=20 OPS=fprop,dgrad,wgrad
    ONLY=1 python tools/layer_probe.py nin2_1          one launch of each requested op (for `ncu -k regex:tc_conv`)

Prints one line per (layer, op): microseconds, algorithmic TFLOP/s, effective GB/s of the compulsory traffic.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convnet_b200 import conv_gemm as cg  # noqa: E402
from convnet_b200 import lib  # noqa: E402
from convnet_b200.abi import GetConvDesc, num_modules  # noqa: E402
from convnet_b200.matrix import CUDAMatrix  # noqa: E402

LAYERS = {
    # name: W, H, Cin, Cout, ky, kx, sy, sx, py, px
    "conv1": (224, 224, 3, 96, 7, 7, 2, 2, 1, 1),
    "conv2": (55, 55, 96, 256, 5, 5, 2, 2, 1, 1),
    "nin2_1": (27, 27, 256, 256, 1, 1, 1, 1, 0, 0),
    "conv3": (14, 14, 256, 384, 3, 3, 1, 1, 1, 1),
    "nin3_1": (14, 14, 384, 768, 1, 1, 1, 1, 0, 0),
    "conv4": (14, 14, 768, 384, 3, 3, 1, 1, 1, 1),
    "nin4_1": (14, 14, 384, 768, 1, 1, 1, 1, 0, 0),
    "nin4_2": (14, 14, 768, 384, 1, 1, 1, 1, 0, 0),
    "conv5": (14, 14, 384, 512, 3, 3, 1, 1, 0, 0),
    "nin5_1": (12, 12, 512, 1024, 1, 1, 1, 1, 0, 0),
    "nin5_2": (12, 12, 1024, 512, 1, 1, 1, 1, 0, 0),
    "fc6": (1, 1, 18432, 4096, 1, 1, 1, 1, 0, 0),
    "fc7": (1, 1, 4096, 4096, 1, 1, 1, 1, 0, 0),
    "fc8": (1, 1, 4096, 1000, 1, 1, 1, 1, 0, 0),
}


def timed(fn, iters, flush):
    fn(); torch.cuda.synchronize()
    if iters <= 1:
        return 0.0
    t0 = time.time()
    while time.time() - t0 < 0.03:
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        if flush is not None:
            flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    names = [a for a in sys.argv[1:] if a in LAYERS] or list(LAYERS)
    N = int(os.environ.get("BATCH", "128"))
    mode = os.environ.get("MODE", "bf16")
    iters = 1 if os.environ.get("ONLY") else int(os.environ.get("ITERS", "20"))
    ops = os.environ.get("OPS", "fprop,dgrad,wgrad").split(",")
    fuse = os.environ.get("FUSE", "1") == "1"
    L = lib.load()
    lib.set_precision(mode)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if os.environ.get("FLUSH", "1") == "1" else None
    total = {o: 0.0 for o in ops}
    for name in names:
        W, H, Cin, Cout, ky, kx, sy, sx, py, px = LAYERS[name]
        modY, modX = num_modules(H, ky, sy, py), num_modules(W, kx, sx, px)
        d = GetConvDesc(Cin, Cout, ky, kx, sy, sx, py, px)
        ish, fsh, tsh = (N, W, H, Cin), (Cout, kx, ky, Cin), (N, modX, modY, Cout)
        K = kx * ky * Cin
        g = torch.Generator(device="cuda").manual_seed(0)
        x = CUDAMatrix(N, W * H * Cin, ish); x.storage.normal_(generator=g)
        w = CUDAMatrix(Cout, K, fsh); w.storage.normal_(generator=g).mul_(1 / np.sqrt(K))
        dv = CUDAMatrix(N, modX * modY * Cout, tsh); dv.storage.normal_(generator=g)
        bias = torch.randn(Cout, device="cuda")
        up = CUDAMatrix(N, modX * modY * Cout, tsh); dn = CUDAMatrix(N, W * H * Cin, ish); dw = CUDAMatrix(Cout, K, fsh)
        if mode == "bf16":
            for m in (x, w, dv):
                L.convnet_b200_bf16_stage(m.ptr, m.storage.numel())
        flops = 2.0 * N * modX * modY * Cout * K

        def f_up():
            if fuse:
                L.convnet_b200_fuse_next(bias.data_ptr(), 1, None)
            cg.convUp(x, w, up, d)

        def f_dn():
            if fuse:
                L.convnet_b200_fuse_next(None, 0, x.ptr)
            cg.convDown(dv, w, dn, d)

        def f_dw():
            cg.convOutp(x, dv, dw, d, 0, 1.0 / N)
        esz = 2 if mode == "bf16" else 4
        byts = {"fprop": x.storage.numel() * esz + w.storage.numel() * esz + up.storage.numel() * 4,
                "dgrad": dv.storage.numel() * esz + w.storage.numel() * esz + dn.storage.numel() * (8 if fuse else 4),
                "wgrad": x.storage.numel() * esz + dv.storage.numel() * esz + dw.storage.numel() * 4}
        for op, fn in (("fprop", f_up), ("dgrad", f_dn), ("wgrad", f_dw)):
            if op not in ops or (op == "dgrad" and name == "conv1"):
                continue
            ms = timed(fn, iters, flush)
            path = lib.last_conv_path()
            total[op] += ms
            if iters > 1:
                print("%-7s b%-4d %-5s %-13s %8.1f us  %7.1f TF/s  %6.0f GB/s" % (
                    name, N, op, path, ms * 1e3, flops / ms / 1e9, byts[op] / ms / 1e6), flush=True)
        L.convnet_b200_bf16_invalidate(None)
    if iters > 1:
        print("TOTAL " + "  ".join("%s %.1f us" % (o, total[o] * 1e3) for o in ops), flush=True)


if __name__ == "__main__":
    main()
