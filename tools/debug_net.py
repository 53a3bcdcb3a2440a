import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from convnet_b200 import lib, net
from test_gpu_net import _torch_net
lib.load(); net.load_host()
MODEL = sys.argv[1] if len(sys.argv) > 1 else "tiny"
for mode in ("fp32", "tf32"):
    lib.set_precision(mode)
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    n = net.Net(MODEL, batch, seed=7)
    g = torch.Generator(device="cuda").manual_seed(11)
    n.input_tensor().normal_(generator=g)
    n.labels_tensor().copy_(torch.randint(0, n.num_classes, (batch,), device="cuda", generator=g, dtype=torch.int32))
    n.fprop(False); n.bprop()
    loss = n.loss()
    ref_loss, params = _torch_net(torch, n, batch, MODEL)
    print(mode, "loss", loss, ref_loss)
    G = n.grads_tensor().double(); edges = n.edges()
    for i, (w, b, K) in params.items():
        off = edges[i][2]; cout = b.shape[0]
        gw = G[off:off + cout * K].view(K, cout)
        if w.dim() == 4: gw = gw.view(w.shape[1], w.shape[2], w.shape[3], cout).permute(3, 0, 1, 2)
        gb = G[off + cout * K:off + cout * K + cout]
        for name, mine, ref in (("w", gw, w.grad / batch), ("b", gb, b.grad / batch)):
            err = ((mine - ref).abs().max() / ref.abs().mean().clamp_min(1e-12)).item()
            print("  ", edges[i][0], name, "err %.3e" % err, "mean|ref| %.3e" % ref.abs().mean().item())
    n.close()
