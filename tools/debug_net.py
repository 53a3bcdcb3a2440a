import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from convnet_b200 import lib, net
from test_gpu_net import _torch_tiny_net
lib.load(); net.load_host()
for mode in ("fp32", "tf32"):
    lib.set_precision(mode)
    batch = 32
    n = net.Net("tiny", batch, seed=7)
    g = torch.Generator(device="cuda").manual_seed(11)
    n.input_tensor().normal_(generator=g)
    n.labels_tensor().copy_(torch.randint(0, 10, (batch,), device="cuda", generator=g, dtype=torch.int32))
    n.fprop(False); n.bprop()
    loss = n.loss()
    ref_loss, params = _torch_tiny_net(torch, n, batch)
    print(mode, "loss", loss, ref_loss)
    G = n.grads_tensor().double(); edges = n.edges()
    for i, (w, b, K) in params.items():
        off = edges[i][2]; cout = w.shape[0]
        gw = G[off:off + cout * K].view(K, cout).view(w.shape[1], w.shape[2], w.shape[3], cout).permute(3, 0, 1, 2)
        gb = G[off + cout * K:off + cout * K + cout]
        for name, mine, ref in (("w", gw, w.grad / batch), ("b", gb, b.grad / batch)):
            err = ((mine - ref).abs().max() / ref.abs().mean().clamp_min(1e-12)).item()
            print("  ", edges[i][0], name, "err %.3e" % err, "mean|ref| %.3e" % ref.abs().mean().item())
    n.close()
lib.set_precision("fp32")
n = net.Net("tiny", 16, seed=3, grad_checker=True)
print("gradcheck fp32", n.grad_check(seed=5)); n.close()
lib.set_precision("tf32")
for model, batch, classes in (("tiny", 32, 10), ("lenet", 100, 10)):
    n = net.Net(model, batch, seed=1)
    g = torch.Generator(device="cuda").manual_seed(0)
    n.input_tensor().normal_(generator=g)
    n.labels_tensor().copy_(torch.randint(0, classes, (batch,), device="cuda", generator=g, dtype=torch.int32))
    losses = [n.train_step(True) / batch for _ in range(60)]
    print(model, ["%.3f" % v for v in losses[::6]])
    n.close()
n = net.Net("c3d", 4, seed=1)
n.input_tensor().normal_(); n.labels_tensor().zero_()
print("c3d loss", n.train_step(True) / 4, math.log(101), torch.isfinite(n.grads_tensor()).all().item())
