"""CPU arm of the bench: the reference's own CPU implementation of the hot path
(eigenmat/cpumat_conv.cc through oracle/_ref/libeigenmat_ref.so; the C restatement
oracle/libconv_oracle.so when _ref is not there) running the AlexNet training step's conv / 1x1 / fc
work — fprop, dgrad (not into the input layer) and wgrad of every weighted edge of
CLS_net_20140801232522 — on the host cores.  Pooling / response-norm / elementwise steps are
<0.1 % of the CPU time and are left out (stated in the `sample` field).

TEST/BENCH INFRASTRUCTURE: this is the baseline being measured NEXT TO the product, never part of it.
The reference's conv path is single-threaded (naive sgemm), so `cores` worker
processes each push their own images through it: that is "all the host threads it can use".
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# (name, W, H, Cin, Cout, k, stride, pad) — SURVEY.md Appendix B; fc layers are 1x1 convs on a 1x1 image
ALEXNET_WEIGHTED = [
    ("conv1", 224, 224, 3, 96, 7, 2, 1), ("conv2", 55, 55, 96, 256, 5, 2, 1), ("nin2_1", 27, 27, 256, 256, 1, 1, 0),
    ("conv3", 14, 14, 256, 384, 3, 1, 1), ("nin3_1", 14, 14, 384, 768, 1, 1, 0), ("conv4", 14, 14, 768, 384, 3, 1, 1),
    ("nin4_1", 14, 14, 384, 768, 1, 1, 0), ("nin4_2", 14, 14, 768, 384, 1, 1, 0), ("conv5", 14, 14, 384, 512, 3, 1, 0),
    ("nin5_1", 12, 12, 512, 1024, 1, 1, 0), ("nin5_2", 12, 12, 1024, 512, 1, 1, 0),
    ("fc6", 1, 1, 18432, 4096, 1, 1, 0), ("fc7", 1, 1, 4096, 4096, 1, 1, 0), ("fc8", 1, 1, 4096, 1000, 1, 1, 0),
]


def _backend():
    from oracle_lib import Oracle, RefLib
    if RefLib.available():
        return RefLib(), "reference"
    return Oracle(), "port"


# Bounded samples, smallest first.  A full step is ~30 s per image on an idle core but minutes when 128 worker processes
# share the memory system, so the sample is chosen per box from a calibration run to fit a time budget.
SAMPLES = (("nin2_1",), ("conv3", "nin3_1", "fc7"), None)      # None = all 14 weighted edges
NOMINAL_S_PER_GFLOP = 2.3        # seconds per GFLOP on one core while all cores run the reference's naive sgemm (128-core box)


def flops_per_image(layers=None):
    f = 0.0
    for i, (name, W, H, Cin, Cout, k, s, p) in enumerate(ALEXNET_WEIGHTED):
        if layers is not None and name not in layers:
            continue
        mod = (W + 2 * p - k) // s + 1
        f += 2.0 * mod * mod * Cout * k * k * Cin * (2 if i == 0 else 3)
    return f


def _fill(r, rows, cols, scale=1.0):
    """pseudo-random Fortran-ordered float32 matrix, cheap for the 75 M-element fc6 weights (a repeated 64 K tile)"""
    tile = (r.standard_normal(65536) * scale).astype(np.float32)
    return np.asfortranarray(np.resize(tile, rows * cols).reshape((rows, cols), order="F"))


def run_step(n_images, seed=0, layers=None):
    """one AlexNet training step's conv/fc work for a batch of n_images on ONE core; returns the seconds spent inside
    the reference's conv calls (input generation and allocation are not timed)."""
    from convnet_b200.abi import GetConvDesc
    lib, _ = _backend()
    r = np.random.RandomState(seed)
    spent = 0.0
    for i, (name, W, H, Cin, Cout, k, s, p) in enumerate(ALEXNET_WEIGHTED):
        if layers is not None and name not in layers:
            continue
        mod = (W + 2 * p - k) // s + 1
        d = GetConvDesc(Cin, Cout, k, k, s, s, p, p)
        ish, fsh, tsh = (n_images, W, H, Cin), (Cout, k, k, Cin), (n_images, mod, mod, Cout)
        x = _fill(r, n_images, W * H * Cin)
        w = _fill(r, Cout, k * k * Cin, 0.01)
        dy = _fill(r, n_images, mod * mod * Cout)
        y = np.zeros((n_images, mod * mod * Cout), dtype=np.float32, order="F")
        dw = np.zeros_like(w)
        dx = np.zeros_like(x) if i > 0 else None
        t0 = time.perf_counter()
        lib.convUp(x, w, y, ish, fsh, tsh, d, 0.0, 1.0)
        lib.convOutp(x, dy, dw, ish, tsh, fsh, d, 0.0, 1.0 / n_images)
        if i > 0:
            lib.convDown(dy, w, dx, tsh, fsh, ish, d, 0.0, 1.0)
        spent += time.perf_counter() - t0
    return spent


def _worker(args):
    n_images, seed, layers = args
    os.environ["OMP_NUM_THREADS"] = "1"
    return run_step(n_images, seed, layers)


class Pool:
    """worker processes kept alive across bench steps"""

    def __init__(self, cores=None):
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores)
        _, self.kind = _backend()

    def close(self):
        self.pool.close(); self.pool.join()

    def _run(self, layers, images_per_core=1):
        t0 = time.perf_counter()
        per = self.pool.map(_worker, [(images_per_core, 100 + i, layers) for i in range(self.cores)])
        return max(per), time.perf_counter() - t0

    def step(self, budget_s=25.0, images_per_core=1):
        """one bounded sample: every core pushes images_per_core image(s) through the largest layer set of SAMPLES whose
        NOMINAL time fits budget_s.  The choice depends on the budget only (i.e. on the command line), never on how busy
        the box happens to be: the nominal cost is NOMINAL_S_PER_GFLOP seconds per GFLOP per core with every core busy
        (measured on the 128-core bench boxes in round 1; an 8-core box is ~3x faster per core).  images/s of the whole
        step is extrapolated by FLOPs when the set is not the full one.  Returns (images_per_second, description)."""
        layers = SAMPLES[0]
        for cand in SAMPLES[1:]:
            if NOMINAL_S_PER_GFLOP * flops_per_image(cand) / 1e9 * images_per_core <= budget_s:
                layers = cand
        slow, wall = self._run(layers, images_per_core)
        images = images_per_core * self.cores
        frac = flops_per_image(layers) / flops_per_image(None)
        value = images * frac / slow           # every core finishes its share within `slow` seconds
        what = "all 14 weighted edges" if layers is None else \
            "layers %s (%.1f%% of the step's FLOPs, images/s extrapolated by FLOPs; the set is fixed by the per-step budget of " \
            "%.0f s at a nominal %.1f s/GFLOP/core, not by box load)" % ("+".join(layers), 100 * frac, budget_s, NOMINAL_S_PER_GFLOP)
        desc = ("AlexNet (CLS_net_20140801232522) training step on the host CPU: conv/1x1/fc fprop+dgrad+wgrad of %s, "
                "%.2f GFLOP/image, %d image(s)/core x %d cores; pool/rnorm/elementwise (<0.1%% of CPU time) omitted; "
                "slowest core %.1f s inside the reference's conv calls, wall %.1f s" % (
                    what, flops_per_image(None) / 1e9, images_per_core, self.cores, slow, wall))
        return value, desc


def measure(budget_s=25.0, cores=None):
    p = Pool(cores)
    try:
        value, desc = p.step(budget_s)
        return {"value": value, "unit": "images/s", "cores": p.cores, "kind": p.kind, "sample": desc}
    finally:
        p.close()


if __name__ == "__main__":
    print(measure(float(sys.argv[1]) if len(sys.argv) > 1 else 25.0, int(sys.argv[2]) if len(sys.argv) > 2 else None))
