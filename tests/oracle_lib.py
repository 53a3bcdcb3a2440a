"""ctypes doors onto oracle/ (TEST INFRASTRUCTURE: the CPU checker).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
`Oracle` wraps oracle/libconv_oracle.so (the plain-C restatement); `RefLib` wraps
oracle/_ref/libeigenmat_ref.so (the reference's own CPU library, when built).
All arrays are numpy float32, Fortran-ordered (rows = images): the reference layout.
"""
import ctypes as ct
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from convnet_b200.abi import ConvDesc, Shape4D  # noqa: E402

ORACLE_DIR = os.path.join(ROOT, "oracle")
FP = ct.POINTER(ct.c_float)
SP = ct.POINTER(Shape4D)


def build_oracle():
    """make -C oracle (C restatement always; _ref when /root/reference is present)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "--no-print-directory"], check=True,
                   stdout=subprocess.DEVNULL)


def _p(a):
    assert a.dtype == np.float32 and a.flags.f_contiguous, (a.dtype, a.flags)
    return a.ctypes.data_as(FP)


def fmat(rows, cols, fill=None, rng=None, kind="randn"):
    """Column-major float32 matrix like a cudamat (rows x cols)."""
    if rng is not None:
        a = rng.standard_normal((rows, cols)) if kind == "randn" else rng.random((rows, cols))
        return np.asfortranarray(a.astype(np.float32))
    a = np.zeros((rows, cols), dtype=np.float32, order="F")
    if fill is not None:
        a[...] = fill
    return a


def s4(t):
    return Shape4D.of(*t)


class Oracle:
    def __init__(self):
        path = os.path.join(ORACLE_DIR, "libconv_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = L = ct.CDLL(path)
        conv_args = [FP, FP, FP, SP, SP, SP, ConvDesc, ct.c_float, ct.c_float, ct.c_int]
        for name in ("oracle_convUp", "oracle_convDown", "oracle_convOutp"):
            getattr(L, name).argtypes = conv_args
            getattr(L, name).restype = None
        L.oracle_convOutpPartial.argtypes = [FP, FP, FP, SP, SP, SP, ConvDesc, ct.c_int, ct.c_int,
                                            ct.c_float, ct.c_float]
        L.oracle_convUp3D.argtypes = [FP, FP, FP, SP, SP, SP, ConvDesc, ct.c_float]
        L.oracle_convDown3D.argtypes = [FP, FP, FP, SP, SP, SP, ConvDesc, ct.c_float]
        L.oracle_convOutp3D.argtypes = [FP, FP, FP, SP, SP, SP, ConvDesc, ct.c_float, ct.c_float]
        L.oracle_pool.argtypes = [ct.c_int, FP, FP, SP, SP, ConvDesc, ct.c_float]
        L.oracle_maxPoolUndo.argtypes = [FP, FP, FP, FP, SP, SP, ConvDesc, ct.c_float]
        L.oracle_avgPoolUndo.argtypes = [FP, FP, SP, SP, ConvDesc, ct.c_float, ct.c_float]
        L.oracle_rnorm.argtypes = [FP, FP, ct.c_long, ct.c_int, ct.c_int, ct.c_float, ct.c_float, ct.c_int]
        L.oracle_rnormUndo.argtypes = [FP, FP, FP, ct.c_long, ct.c_int, ct.c_int, ct.c_float,
                                       ct.c_float, ct.c_int]
        L.oracle_extract_patches.argtypes = [FP, FP, FP, FP, FP] + [ct.c_int] * 6
        L.oracle_extract_patches.restype = ct.c_int

    # --- conv (2-D) ---
    def convUp(self, images, filters, targets, ish, fsh, tsh, d, scaleTargets=0.0, scaleOutput=1.0, conv=True):
        self.lib.oracle_convUp(_p(images), _p(filters), _p(targets), s4(ish), s4(fsh), s4(tsh), d,
                               scaleTargets, scaleOutput, int(conv))

    def convDown(self, derivs, filters, targets, dsh, fsh, tsh, d, scaleTargets=0.0, scaleOutput=1.0, conv=True):
        self.lib.oracle_convDown(_p(derivs), _p(filters), _p(targets), s4(dsh), s4(fsh), s4(tsh), d,
                                 scaleTargets, scaleOutput, int(conv))

    def convOutp(self, images, derivs, targets, ish, dsh, tsh, d, scaleTargets=0.0, scaleOutput=1.0, conv=True):
        self.lib.oracle_convOutp(_p(images), _p(derivs), _p(targets), s4(ish), s4(dsh), s4(tsh), d,
                                 scaleTargets, scaleOutput, int(conv))

    def convOutpPartial(self, images, derivs, targets, ish, dsh, tsh, d, psy, psx, scaleTargets=0.0,
                        scaleOutput=1.0):
        self.lib.oracle_convOutpPartial(_p(images), _p(derivs), _p(targets), s4(ish), s4(dsh), s4(tsh), d,
                                        psy, psx, scaleTargets, scaleOutput)

    # --- conv (3-D) ---
    def convUp3D(self, images, filters, targets, ish, fsh, tsh, d, scaleTargets=0.0):
        self.lib.oracle_convUp3D(_p(images), _p(filters), _p(targets), s4(ish), s4(fsh), s4(tsh), d, scaleTargets)

    def convDown3D(self, derivs, filters, targets, dsh, fsh, tsh, d, scaleTargets=0.0):
        self.lib.oracle_convDown3D(_p(derivs), _p(filters), _p(targets), s4(dsh), s4(fsh), s4(tsh), d, scaleTargets)

    def convOutp3D(self, images, derivs, targets, ish, dsh, tsh, d, scaleTargets=0.0, scaleOutput=1.0):
        self.lib.oracle_convOutp3D(_p(images), _p(derivs), _p(targets), s4(ish), s4(dsh), s4(tsh), d,
                                   scaleTargets, scaleOutput)

    # --- pooling ---
    def pool(self, is_max, images, targets, ish, tsh, d, scaleOutput=1.0):
        self.lib.oracle_pool(int(is_max), _p(images), _p(targets), s4(ish), s4(tsh), d, scaleOutput)

    def maxPoolUndo(self, images, maxGrads, maxActs, targets, ish, gsh, d, scaleTargets=0.0):
        self.lib.oracle_maxPoolUndo(_p(images), _p(maxGrads), _p(maxActs), _p(targets), s4(ish), s4(gsh), d,
                                    scaleTargets)

    def avgPoolUndo(self, avgGrads, targets, gsh, tsh, d, scaleTargets=0.0, scaleOutput=1.0):
        self.lib.oracle_avgPoolUndo(_p(avgGrads), _p(targets), s4(gsh), s4(tsh), d, scaleTargets, scaleOutput)

    # --- input pipeline ---
    def extract_patches(self, images, patches, width_offset, height_offset, flip, N, C, W, H, pw, ph):
        """images: flat float32, image n at [n*C*H*W, (n+1)*C*H*W) as (c, y, x); patches: flat, image fastest"""
        return self.lib.oracle_extract_patches(_p(images), _p(patches), _p(width_offset), _p(height_offset), _p(flip),
                                               N, C, W, H, pw, ph)

    # --- response norm ---
    def rnorm(self, images, targets, numFilters, sizeF, addScale, powScale, blocked=False):
        self.lib.oracle_rnorm(_p(images), _p(targets), images.size, numFilters, sizeF, addScale, powScale,
                              int(blocked))

    def rnormUndo(self, outGrads, inputs, targets, numFilters, sizeF, addScale, powScale, blocked=False):
        self.lib.oracle_rnormUndo(_p(outGrads), _p(inputs), _p(targets), inputs.size, numFilters, sizeF,
                                  addScale, powScale, int(blocked))


class RefLib:
    """The reference's own CPU library (oracle/_ref/libeigenmat_ref.so)."""

    PATH = os.path.join(ORACLE_DIR, "_ref", "libeigenmat_ref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = L = ct.CDLL(self.PATH)
        IP = ct.POINTER(ct.c_int)
        self._IP = IP
        for name in ("ref_convUp", "ref_convDown", "ref_convOutp"):
            getattr(L, name).argtypes = [FP, FP, FP, IP, IP, IP, ConvDesc, ct.c_float, ct.c_float, ct.c_int]
            getattr(L, name).restype = None
        L.ref_rnorm.argtypes = [FP, FP, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_float, ct.c_float, ct.c_int]
        L.ref_rnormUndo.argtypes = [FP, FP, FP, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_float, ct.c_float,
                                    ct.c_int]
        if hasattr(L, "ref_extract_patches"):                 # a prebuilt _ref from before this door existed lacks it
            L.ref_extract_patches.argtypes = [FP, FP, FP, FP, FP] + [ct.c_int] * 6
            L.ref_extract_patches.restype = ct.c_int

    def _s(self, t):
        return (ct.c_int * 4)(*[int(v) for v in t])

    def convUp(self, images, filters, targets, ish, fsh, tsh, d, scaleTargets=0.0, scaleOutput=1.0, conv=True):
        self.lib.ref_convUp(_p(images), _p(filters), _p(targets), self._s(ish), self._s(fsh), self._s(tsh), d,
                            scaleTargets, scaleOutput, int(conv))

    def convDown(self, derivs, filters, targets, dsh, fsh, tsh, d, scaleTargets=0.0, scaleOutput=1.0, conv=True):
        self.lib.ref_convDown(_p(derivs), _p(filters), _p(targets), self._s(dsh), self._s(fsh), self._s(tsh), d,
                              scaleTargets, scaleOutput, int(conv))

    def convOutp(self, images, derivs, targets, ish, dsh, tsh, d, scaleTargets=0.0, scaleOutput=1.0, conv=True):
        self.lib.ref_convOutp(_p(images), _p(derivs), _p(targets), self._s(ish), self._s(dsh), self._s(tsh), d,
                              scaleTargets, scaleOutput, int(conv))

    def extract_patches(self, images, patches, width_offset, height_offset, flip, N, C, W, H, pw, ph):
        return self.lib.ref_extract_patches(_p(images), _p(patches), _p(width_offset), _p(height_offset), _p(flip),
                                            N, C, W, H, pw, ph)

    def rnorm(self, images, targets, numFilters, sizeF, addScale, powScale, blocked=False):
        self.lib.ref_rnorm(_p(images), _p(targets), images.shape[0], images.shape[1], numFilters, sizeF,
                           addScale, powScale, int(blocked))

    def rnormUndo(self, outGrads, inputs, targets, numFilters, sizeF, addScale, powScale, blocked=False):
        self.lib.ref_rnormUndo(_p(outGrads), _p(inputs), _p(targets), inputs.shape[0], inputs.shape[1],
                               numFilters, sizeF, addScale, powScale, int(blocked))


def Diff(a, b):
    """py/test_conv.py:382-385 — the reference's own parity metric."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(a + b).mean()
    return float(np.abs(a - b).max() / scale) if scale > 0 else float(np.abs(a - b).max())
