"""Every ABI-2 entry point (the bare names of cudamat/cudamat_conv.cuh:8-78 — what a USE_GEMM_KERNELS=no build of the
reference links) is EXECUTED against the CPU oracle: convUp, convDown, convOutp (with and without partial sums), localUp,
localDown, localOutp, MaxPool, AvgPool (no scale arguments), MaxPoolUndo, AvgPoolUndo, UpSample, DownSample,
ResponseNormCrossMap, ResponseNormCrossMapUndo (takes `acts`), SetupTexture.  A signature slip in csrc/abi.cu's
ABI-2 half shows up here as a wrong result or a crash.  Tolerances as in tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

from cases import F, Z
from convnet_b200.abi import GetConvDesc, num_modules
from oracle_lib import Diff

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "tf32": 5e-3, "bf16": 2.5e-2}
TOL_MEM = 1e-4


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from convnet_b200 import conv_gemm as cg
    from convnet_b200 import lib
    from convnet_b200.matrix import CUDAMatrix
    lib.load()

    class G:
        pass
    g = G()
    g.cc2, g.lib, g.torch = cg.cc2, lib, torch
    g.up = lambda a, s4=None: CUDAMatrix.from_numpy(a, s4)

    def nan_like(r, c, s4=None):
        m = CUDAMatrix(r, c, s4)
        m.fill_(float("nan"))
        return m
    g.nan = nan_like
    yield g
    lib.set_precision("tf32")


CONV = {
    # name: N, W, H, Cin, Cout, ky, kx, sy, sx, py, px
    "testconv_full": (128, 12, 12, 32, 64, 3, 3, 2, 2, 1, 1),     # py/test_conv.py:394-441
    "alex_conv3_b32": (32, 14, 14, 256, 384, 3, 3, 1, 1, 1, 1),
    "ragged": (37, 9, 8, 12, 20, 3, 2, 2, 1, 1, 0),
    "mnist_conv1": (100, 28, 28, 1, 48, 4, 4, 1, 1, 0, 0),
}


def _case(shape, seed=11):
    N, W, H, Cin, Cout, ky, kx, sy, sx, py, px = shape
    modY, modX = num_modules(H, ky, sy, py), num_modules(W, kx, sx, px)
    d = GetConvDesc(Cin, Cout, ky, kx, sy, sx, py, px)
    ish, fsh, tsh = (N, W, H, Cin), (Cout, kx, ky, Cin), (N, modX, modY, Cout)
    r = np.random.RandomState(seed)
    return (d, ish, fsh, tsh, F(r.randn(N, W * H * Cin)), F(r.randn(Cout, kx * ky * Cin) / np.sqrt(kx * ky * Cin)),
            F(r.randn(N, modX * modY * Cout)))


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("case", sorted(CONV))
def test_abi2_conv(gpu, oracle, case, mode):
    gpu.lib.set_precision(mode)
    d, ish, fsh, tsh, images, filters, derivs = _case(CONV[case])
    N = ish[0]
    gi, gf, gd = gpu.up(images, ish), gpu.up(filters, fsh), gpu.up(derivs, tsh)
    gpu.cc2.SetupTexture(gi)                                                     # cudamat_conv.cuh:8 (a no-op here)
    out = gpu.nan(N, derivs.shape[1], tsh); gpu.cc2.convUp(gi, gf, out, d, 0)
    ref = Z(*derivs.shape); oracle.convUp(images, filters, ref, ish, fsh, tsh, d)
    assert Diff(out.asarray(), ref) < TOL[mode]
    init = F(np.random.RandomState(2).randn(*images.shape))
    out = gpu.up(init, ish); gpu.cc2.convDown(gd, gf, out, d, 1)
    ref = init.copy(order="F"); oracle.convDown(derivs, filters, ref, tsh, fsh, ish, d, 1.0)
    assert Diff(out.asarray(), ref) < TOL[mode]
    out = gpu.nan(*filters.shape, fsh); gpu.cc2.convOutp(gi, gd, out, d, 0, 1.0 / N)        # partialSum 0/0
    ref = Z(*filters.shape); oracle.convOutp(images, derivs, ref, ish, tsh, fsh, d, 0.0, 1.0 / N)
    assert Diff(out.asarray(), ref) < TOL[mode]


def test_abi2_conv_outp_partial_sums(gpu, oracle):
    """convOutp with partialSumY/X (weightacts.cu:3126-3170): one [Cout x K] block per module rectangle."""
    gpu.lib.set_precision("fp32")
    d, ish, fsh, tsh, images, filters, derivs = _case((32, 12, 12, 8, 16, 3, 3, 1, 1, 1, 1))
    psy, psx = 4, 6
    chunks = (12 // psy) * (12 // psx)
    Cout, K = filters.shape
    gi, gd = gpu.up(images, ish), gpu.up(derivs, tsh)
    out = gpu.nan(Cout, K * chunks, (Cout, 3, 3, 8 * chunks)); gpu.cc2.convOutpPartial(gi, gd, out, d, psy, psx, 0, 1.0)
    ref = Z(Cout, K * chunks); oracle.convOutpPartial(images, derivs, ref, ish, tsh, (Cout, 3, 3, 8 * chunks), d, psy, psx)
    assert Diff(out.asarray(), ref) < 1e-4
    # the blocks add up to the full sum over modules
    full = Z(Cout, K); oracle.convOutp(images, derivs, full, ish, tsh, fsh, d)
    assert Diff(out.asarray().reshape(Cout, chunks, K).sum(1), full) < 1e-4


def test_abi2_local(gpu, oracle):
    N, W, H, Cin, Cout = 8, 6, 6, 4, 8
    d = GetConvDesc(Cin, Cout, 3, 3, 1, 1, 1, 1)
    mods = 36
    ish, fsh, tsh = (N, W, H, Cin), (Cout, 3, 3, Cin * mods), (N, 6, 6, Cout)
    r = np.random.RandomState(4)
    images, filters, derivs = F(r.randn(N, W * H * Cin)), F(r.randn(Cout, 9 * Cin * mods)), F(r.randn(N, mods * Cout))
    gi, gf, gd = gpu.up(images, ish), gpu.up(filters, fsh), gpu.up(derivs, tsh)
    out = gpu.nan(N, mods * Cout, tsh); gpu.cc2.localUp(gi, gf, out, d)
    ref = Z(N, mods * Cout); oracle.convUp(images, filters, ref, ish, fsh, tsh, d, conv=False)
    assert Diff(out.asarray(), ref) < 1e-4
    out = gpu.nan(*images.shape, ish); gpu.cc2.localDown(gd, gf, out, d)
    ref = Z(*images.shape); oracle.convDown(derivs, filters, ref, tsh, fsh, ish, d, conv=False)
    assert Diff(out.asarray(), ref) < 1e-4
    out = gpu.nan(*filters.shape, fsh); gpu.cc2.localOutp(gi, gd, out, d)
    ref = Z(*filters.shape); oracle.convOutp(images, derivs, ref, ish, tsh, fsh, d, conv=False)
    assert Diff(out.asarray(), ref) < 1e-4


POOL = {
    # N, W, H, C, k, s, p
    "alex_pool1": (32, 110, 110, 8, 3, 2, 1),
    "alex_pool5": (64, 12, 12, 64, 3, 2, 1),
    "mnist_pool": (100, 25, 25, 6, 4, 2, 0),
    "ragged": (7, 9, 9, 5, 3, 2, 1),
}


@pytest.mark.parametrize("case", sorted(POOL))
def test_abi2_pool(gpu, oracle, case):
    N, W, H, C, k, s, p = POOL[case]
    mod = num_modules(W, k, s, p)
    d = GetConvDesc(C, C, k, k, s, s, p, p)
    ish, psh = (N, W, H, C), (N, mod, mod, C)
    r = np.random.RandomState(8)
    images = F(np.round(r.rand(N, W * H * C) * 8) / 8)          # quantised: ties
    grads = F(r.randn(N, mod * mod * C))
    gi, gg = gpu.up(images, ish), gpu.up(grads, psh)
    mx = gpu.nan(N, mod * mod * C, psh); gpu.cc2.MaxPool(gi, mx, d)
    rmx = Z(N, mod * mod * C); oracle.pool(True, images, rmx, ish, psh, d)
    assert np.array_equal(mx.asarray(), rmx)
    av = gpu.nan(N, mod * mod * C, psh); gpu.cc2.AvgPool(gi, av, d)
    rav = Z(N, mod * mod * C); oracle.pool(False, images, rav, ish, psh, d)
    assert Diff(av.asarray(), rav) < TOL_MEM
    for st in (0.0, 1.0):
        init = F(r.randn(*images.shape))
        out = gpu.up(init, ish) if st else gpu.nan(*images.shape, ish)
        gpu.cc2.MaxPoolUndo(gi, gg, mx, out, d, st)
        ref = init.copy(order="F"); oracle.maxPoolUndo(images, grads, rmx, ref, ish, psh, d, st)
        assert Diff(out.asarray(), ref) < TOL_MEM
        out = gpu.up(init, ish) if st else gpu.nan(*images.shape, ish)
        gpu.cc2.AvgPoolUndo(gg, out, d, st)
        ref = init.copy(order="F"); oracle.avgPoolUndo(grads, ref, psh, ish, d, st)
        assert Diff(out.asarray(), ref) < TOL_MEM


def test_abi2_up_down_sample(gpu, oracle):
    N, W, C, f = 12, 5, 6, 3
    small, big = (N, W, W, C), (N, W * f, W * f, C)
    d = GetConvDesc(C, C, f, f, f, f, 0, 0)
    r = np.random.RandomState(1)
    a = F(r.randn(N, W * W * C)); ga = gpu.up(a, small)
    out = gpu.nan(N, W * f * W * f * C, big); gpu.cc2.UpSample(ga, out, f)
    ref = Z(N, W * f * W * f * C); oracle.avgPoolUndo(a, ref, small, big, d, 0.0, float(f * f))
    assert Diff(out.asarray(), ref) < TOL_MEM
    b = F(r.randn(N, W * f * W * f * C)); gb = gpu.up(b, big)
    out = gpu.nan(N, W * W * C, small); gpu.cc2.DownSample(gb, out, f)
    ref = Z(N, W * W * C); oracle.pool(False, b, ref, big, small, d)
    assert Diff(out.asarray(), ref) < TOL_MEM


RNORM = {
    # N, W, H, F, sizeF, alpha, beta
    "alex_rnorm1": (16, 55, 55, 96, 24, 5e-4, 0.75),
    "alex_rnorm2": (32, 14, 14, 256, 64, 5e-4, 0.75),
    "testconv": (128, 12, 12, 32, 8, 0.005, 0.75),
    "even_window": (5, 3, 3, 10, 4, 0.1, 1.0),
}


@pytest.mark.parametrize("blocked", [False, True])
@pytest.mark.parametrize("case", sorted(RNORM))
def test_abi2_rnorm(gpu, oracle, case, blocked):
    N, W, H, nf, sizeF, alpha, beta = RNORM[case]
    ish = (N, W, H, nf)
    r = np.random.RandomState(6)
    x, dy = F(r.randn(N, W * H * nf)), F(r.randn(N, W * H * nf))
    gx, gdy = gpu.up(x, ish), gpu.up(dy, ish)
    acts = gpu.nan(*x.shape, ish); gpu.cc2.ResponseNormCrossMap(gx, acts, sizeF, alpha, beta, blocked)
    ref = Z(*x.shape); oracle.rnorm(x, ref, nf, sizeF, alpha, beta, blocked)
    assert Diff(acts.asarray(), ref) < TOL_MEM
    # ABI-2 passes the forward output as `acts` (response_norm_edge.cc:53-66 -> Matrix::ConvResponseNormCrossMapUndo)
    out = gpu.nan(*x.shape, ish); gpu.cc2.ResponseNormCrossMapUndo(gdy, gx, out, sizeF, alpha, beta, blocked, acts=acts)
    ref = Z(*x.shape); oracle.rnormUndo(dy, x, ref, nf, sizeF, alpha, beta, blocked)
    assert Diff(out.asarray(), ref) < 2 * TOL_MEM
