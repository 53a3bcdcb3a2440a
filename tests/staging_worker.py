"""Worker of tests/test_gpu_staging.py: runs with CONVNET_B200_STAGE_VERIFY=1, where every use of a staged bf16 copy
re-converts the fp32 source and aborts on a mismatch.
  mode "train": training steps of a net in bf16 mode — every bf16 copy written by a producing kernel (conv / pool-undo /
                rnorm epilogues, dropout, SGD) must be bit-identical to a conversion of the fp32 tensor it shadows;
  mode "params": (no verify mode needed) training steps with fixed seeds, then a checksum of the parameter bits — the test
                compares runs whose environment switches fusions on and off;
  mode "stale": stage a tensor, overwrite it behind the library's back, use it — the library must notice and abort."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from convnet_b200 import conv_gemm as cg  # noqa: E402
from convnet_b200 import lib  # noqa: E402
from convnet_b200.abi import GetConvDesc  # noqa: E402
from convnet_b200.matrix import CUDAMatrix  # noqa: E402
from convnet_b200.net import Net  # noqa: E402

mode = sys.argv[1]
assert mode == "params" or os.environ.get("CONVNET_B200_STAGE_VERIFY") == "1"
L = lib.load()
lib.set_precision("bf16")
if mode == "train":
    model, batch, steps = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    n = Net(model, batch, seed=3)
    n.input_tensor().normal_()
    n.labels_tensor().copy_(torch.randint(0, n.num_classes, (batch,), device="cuda", dtype=torch.int32))
    for _ in range(steps):
        loss = n.train_step(True)
        assert np.isfinite(loss), loss
    n.fprop(False)                      # inference pass: dropout off, a different last writer of those layers
    torch.cuda.synchronize()
    n.close()
    print("VERIFY-TRAIN-OK")
elif mode == "params":
    model, batch, steps = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    torch.manual_seed(7)
    n = Net(model, batch, seed=3)
    n.input_tensor().normal_()
    n.labels_tensor().copy_(torch.randint(0, n.num_classes, (batch,), device="cuda", dtype=torch.int32))
    losses = [n.train_step(True) for _ in range(steps)]
    bits = n.params_tensor().view(torch.int32).to(torch.int64)
    print("PARAMS", int((bits * (torch.arange(bits.numel(), device="cuda") % 8191 + 1)).sum().item()), " ".join("%.9g" % v for v in losses))
    n.close()
elif mode == "stale":
    N, W, Cin, Cout = 128, 8, 64, 64
    d = GetConvDesc(Cin, Cout, 3, 3, 1, 1, 1, 1)
    x = CUDAMatrix(N, W * W * Cin, (N, W, W, Cin)); x.storage.normal_()
    w = CUDAMatrix(Cout, 9 * Cin, (Cout, 3, 3, Cin)); w.storage.normal_()
    y = CUDAMatrix(N, W * W * Cout, (N, W, W, Cout))
    L.convnet_b200_bf16_stage(x.ptr, x.storage.numel())
    cg.convUp(x, w, y, d); torch.cuda.synchronize()
    print("FIRST-USE-OK", flush=True)
    x.storage.add_(1.0)                 # a write the library cannot see, and no invalidate
    torch.cuda.synchronize()
    cg.convUp(x, w, y, d); torch.cuda.synchronize()
    print("NOT-DETECTED")
