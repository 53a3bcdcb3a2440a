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
    n = net.Net("tiny", 16, seed=3, grad_checker=True)
    res = n.grad_check(seed=5)
    n.close()
    assert len(res) == 4
    for name, eps, dw, db in res:
        assert dw < 0.01 and db < 0.01, (name, eps, dw, db)


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
    assert ((a - b).abs().max() / (a + b).abs().mean()).item() < 2e-2


@pytest.mark.parametrize("model,batch,classes", [("tiny", 32, 10), ("lenet", 100, 10)])
def test_training_reduces_the_loss(env, model, batch, classes):
    torch, lib, net = env
    lib.set_precision("tf32")
    n = net.Net(model, batch, seed=1)
    g = torch.Generator(device="cuda").manual_seed(0)
    n.input_tensor().normal_(generator=g)
    n.labels_tensor().copy_(torch.randint(0, classes, (batch,), device="cuda", generator=g, dtype=torch.int32))
    losses = [n.train_step(True) / batch for _ in range(60)]
    n.close()
    assert all(math.isfinite(v) for v in losses)
    assert abs(losses[0] - math.log(classes)) < 0.7            # untrained softmax: ~log(#classes)
    assert losses[-1] < 0.5 * losses[0], losses[::10]          # memorises one batch


def test_alexnet_step_small_batch(env):
    torch, lib, net = env
    lib.set_precision("tf32")
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


def test_c3d_video_net_step(env):
    torch, lib, net = env
    lib.set_precision("tf32")
    n = net.Net("c3d", 4, seed=1)
    g = torch.Generator(device="cuda").manual_seed(0)
    n.input_tensor().normal_(generator=g)
    n.labels_tensor().copy_(torch.randint(0, 101, (4,), device="cuda", generator=g, dtype=torch.int32))
    l0 = n.train_step(True) / 4
    assert math.isfinite(l0) and abs(l0 - math.log(101)) < 1.5
    assert torch.isfinite(n.grads_tensor()).all()
    n.close()
