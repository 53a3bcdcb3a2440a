"""GPU tests of the native host layer (Matrix / Edge / ConvNet / GradChecker mirror, convnet_b200/host)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    from convnet_b200 import lib, net
    lib.load(); net.load_host()
    yield torch, lib, net
    lib.set_precision("tf32")


def test_run_grad_check_passes_at_reference_tolerance(env):
    """src/grad_check.cc: mean |analytic-numeric|/scale < 0.01 for weights and biases of every weighted edge
    (conv, 1x1, fc behind pool / rnorm / avg-pool), in the exact-fp32 mode the check is specified for."""
    torch, lib, net = env
    lib.set_precision("fp32")
    for seed in (1, 5, 9):
        n = net.Net("gradcheck", 8, seed=3, grad_checker=True)
        res = n.grad_check(seed=seed)
        n.close()
        assert len(res) == 3
        for name, eps, dw, db in res:
            assert dw < 0.01 and db < 0.01, (seed, res)


def test_analytic_gradients_agree_between_fp32_and_tf32(env):
    torch, lib, net = env
    grads = {}
    for mode in ("fp32", "tf32"):
        lib.set_precision(mode)
        n = net.Net("tiny", 32, seed=3)
        g = torch.Generator(device="cuda").manual_seed(0)
        n.input_tensor().normal_(generator=g)
        n.labels_tensor().copy_(torch.randint(0, 10, (32,), device="cuda", generator=g, dtype=torch.int32))
        n.fprop(False); n.bprop()
        grads[mode] = n.grads_tensor().clone()
        n.close()
    a, b = grads["fp32"].double(), grads["tf32"].double()
    # relative L2: tf32 rounding flips a few ReLU / max-pool decisions in a 32-image batch, which moves individual
    # gradient entries by whole contributions; the aggregate stays within a few percent
    assert ((a - b).norm() / a.norm()).item() < 5e-2


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("model,batch,classes", [("tiny", 32, 10), ("lenet", 100, 10)])
def test_training_reduces_the_loss(env, model, batch, classes, mode):
    torch, lib, net = env
    lib.set_precision(mode)
    n = net.Net(model, batch, seed=1)
    g = torch.Generator(device="cuda").manual_seed(0)
    n.input_tensor().normal_(generator=g)
    n.labels_tensor().copy_(torch.randint(0, classes, (batch,), device="cuda", generator=g, dtype=torch.int32))
    losses = [n.train_step(True) / batch for _ in range(60)]
    n.close()
    assert all(math.isfinite(v) for v in losses)
    assert abs(losses[0] - math.log(classes)) < 1.2            # untrained softmax: ~log(#classes)
    assert min(losses[30:]) < 0.8 * losses[0], losses[::6]     # SGD on one fixed batch drives the loss down


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_alexnet_step_small_batch(env, mode):
    torch, lib, net = env
    lib.set_precision(mode)
    n = net.Net("alexnet", 8, seed=1)
    assert n.num_params == 104321024                             # 104,321,000 + 128-float padding per edge
    assert abs(n.flops_train / 8 / 1e9 - 11.87) < 0.01
    g = torch.Generator(device="cuda").manual_seed(0)
    n.input_tensor().normal_(generator=g)
    n.labels_tensor().copy_(torch.randint(0, 1000, (8,), device="cuda", generator=g, dtype=torch.int32))
    p0 = n.params_tensor().clone()
    l0 = n.train_step(True) / 8
    l1 = n.train_step(True) / 8
    assert math.isfinite(l0) and math.isfinite(l1) and abs(l0 - math.log(1000)) < 1.0
    assert torch.isfinite(n.grads_tensor()).all() and not torch.equal(p0, n.params_tensor())
    out = n.output_tensor().view(1000, 8)
    assert torch.allclose(out.sum(0), torch.ones(8, device="cuda"), atol=1e-4)      # softmax rows
    n.close()


def test_traced_step_reports_an_ordered_timeline(env):
    """ConvNet::TraceStep (cnb_net_trace_step): one real training step with timing events — milestones in order, one SGD end
    per bucket, no exchange window on a single rank — and the step it ran is a normal step (parameters moved)."""
    torch, lib, net = env
    lib.set_precision("bf16")
    n = net.Net("alexnet", 128, seed=1)
    g = torch.Generator(device="cuda").manual_seed(0)
    n.input_tensor().normal_(generator=g)
    n.labels_tensor().copy_(torch.randint(0, 1000, (128,), device="cuda", generator=g, dtype=torch.int32))
    n.train_step(False)
    p0 = n.params_tensor().clone()
    t = n.trace_step()
    assert 0 < t["fprop_end_ms"] < t["bprop_compute_end_ms"] <= t["step_end_ms"] < 100
    assert len(t["buckets"]) >= 2 and abs(sum(b["MB"] for b in t["buckets"]) - 4 * 104321024 / 1e6) < 1.0
    for b in t["buckets"]:
        assert b["exchange_begin_ms"] == -1 and b["exchange_end_ms"] == -1
        assert t["fprop_end_ms"] < b["sgd_end_ms"] <= t["step_end_ms"] + 1e-3
    assert not torch.equal(p0, n.params_tensor())
    l = n.train_step(True)                                       # and the ordinary step still works afterwards
    assert math.isfinite(l)
    n.close()
    lib.set_precision("tf32")


def test_c3d_video_net_step(env):
    torch, lib, net = env
    lib.set_precision("tf32")
    n = net.Net("c3d", 4, seed=1)
    g = torch.Generator(device="cuda").manual_seed(0)
    n.input_tensor().normal_(generator=g)
    n.labels_tensor().copy_(torch.randint(0, 101, (4,), device="cuda", generator=g, dtype=torch.int32))
    l0 = n.train_step(True) / 4
    assert math.isfinite(l0) and l0 < math.log(101) + 5
    assert torch.isfinite(n.grads_tensor()).all()
    n.close()


# float64 PyTorch autograd mirrors of models.cc nets, sharing the native net's parameters.
# spec entries: ("conv", cout, k, stride, pad, relu) | ("maxpool", k, s, p) | ("avgpool", k, s, p) |
#               ("rnorm", k, alpha, beta, relu) | ("fc", cout)
TORCH_NETS = {
    "tiny": dict(cin=8, size=12, spec=[("conv", 16, 3, 1, 1, True), ("maxpool", 3, 2, 1), ("rnorm", 8, 0.01, 0.75, True),
                                      ("conv", 24, 1, 1, 0, True), ("conv", 16, 3, 2, 1, True), ("avgpool", 2, 2, 0),
                                      ("fc", 10)]),
    "gradcheck": dict(cin=4, size=8, spec=[("conv", 8, 3, 1, 1, False), ("avgpool", 3, 2, 1), ("rnorm", 4, 0.01, 0.75, False),
                                          ("conv", 12, 1, 1, 0, False), ("fc", 5)]),
}


def _torch_net(torch, n, batch, model):
    import torch.nn.functional as Fn
    cfg = TORCH_NETS[model]
    P = n.params_tensor().double()
    edges = n.edges()
    params = {}

    def conv_w(i, cout, cin, k):
        off, size = edges[i][2], edges[i][3]
        K = cin * k * k
        flat = P[off:off + size].clone()
        w = flat[:cout * K].view(K, cout).view(cin, k, k, cout).permute(3, 0, 1, 2).contiguous().requires_grad_(True)
        b = flat[cout * K:cout * K + cout].clone().requires_grad_(True)
        params[i] = (w, b, K)
        return w, b

    C, S = cfg["cin"], cfg["size"]
    h = n.input_tensor().double().view(C, S, S, batch).permute(3, 0, 1, 2).contiguous()       # [N, C, H, W]
    labels = n.labels_tensor().long()
    for i, e in enumerate(cfg["spec"]):
        if e[0] == "conv":
            _, cout, k, s, p, relu = e
            w, b = conv_w(i, cout, h.shape[1], k)
            h = Fn.conv2d(h, w, b, stride=s, padding=p)
            h = torch.relu(h) if relu else h
        elif e[0] == "maxpool":
            h = Fn.max_pool2d(h, e[1], e[2], e[3])
        elif e[0] == "avgpool":
            h = Fn.avg_pool2d(h, e[1], e[2], e[3], count_include_pad=False)
        elif e[0] == "rnorm":                                  # window [j - k/2, j - k/2 + k) (gemm.cu:475-477)
            _, k, a, bpow, relu = e
            F_ = h.shape[1]
            sq = Fn.pad(h * h, (0, 0, 0, 0, k // 2, k - k // 2 - 1))
            Ssum = sum(sq[:, j:j + F_] for j in range(k))
            h = h * (1 + a * Ssum) ** (-bpow)
            h = torch.relu(h) if relu else h
        elif e[0] == "fc":                                     # features flattened as x + W*(y + H*c)
            cout = e[1]
            K = h.shape[1] * h.shape[2] * h.shape[3]
            off, size = edges[i][2], edges[i][3]
            flat = P[off:off + size].clone()
            w = flat[:cout * K].view(K, cout).clone().requires_grad_(True)           # [K, cout], K = (c, y, x)
            b = flat[cout * K:cout * K + cout].clone().requires_grad_(True)
            params[i] = (w, b, K)
            h = h.reshape(batch, K) @ w + b
    loss = Fn.cross_entropy(h, labels, reduction="sum")
    loss.backward()
    return loss.item(), params


@pytest.mark.parametrize("model", sorted(TORCH_NETS))
def test_backprop_matches_float64_autograd(env, model):
    """Every backward op of the chain (wgrad, dgrad, bias grad, max/avg-pool undo, response-norm undo, ReLU/softmax
    derivatives) against an independent float64 PyTorch autograd model with the same parameters."""
    torch, lib, net = env
    for mode, tol in (("fp32", 2e-5), ("tf32", 5e-2), ("bf16", 1.5e-1)):
        lib.set_precision(mode)
        batch = 32
        n = net.Net(model, batch, seed=7)
        g = torch.Generator(device="cuda").manual_seed(11)
        n.input_tensor().normal_(generator=g)
        n.labels_tensor().copy_(torch.randint(0, n.num_classes, (batch,), device="cuda", generator=g, dtype=torch.int32))
        n.fprop(False); n.bprop()
        loss = n.loss()
        ref_loss, params = _torch_net(torch, n, batch, model)
        assert abs(loss - ref_loss) / ref_loss < {"fp32": 1e-5, "tf32": 2e-3, "bf16": 1e-2}[mode]
        G = n.grads_tensor().double()
        edges = n.edges()
        for i, (w, b, K) in params.items():
            off = edges[i][2]
            cout = b.shape[0]
            gw = G[off:off + cout * K].view(K, cout)
            if w.dim() == 4:
                gw = gw.view(w.shape[1], w.shape[2], w.shape[3], cout).permute(3, 0, 1, 2)
            gb = G[off + cout * K:off + cout * K + cout]
            for name, mine, ref in (("w", gw, w.grad / batch), ("b", gb, b.grad / batch)):   # scale_gradients / batch
                if mode == "fp32":      # exact arithmetic: every entry matches
                    err = ((mine - ref).abs().max() / ref.abs().mean().clamp_min(1e-12)).item()
                else:                   # tf32: relative L2 (rounding flips a few ReLU / max-pool decisions)
                    err = ((mine - ref).norm() / ref.norm().clamp_min(1e-12)).item()
                assert err < tol, (mode, edges[i][0], name, err)
        n.close()
