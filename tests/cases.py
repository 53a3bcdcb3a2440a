"""Shared test-case plumbing: golden loading, geometry helpers (used by CPU and GPU tests)."""
import os

import numpy as np

from convnet_b200.abi import GetConvDesc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_2D = ["ref2d_testconv", "ref2d_rect", "ref2d_conv1"]
GOLDEN_3D = ["ref3d_testconv", "ref3d_c3d"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {}
    for k in z.files:
        v = z[k]
        g[k] = v.item() if v.ndim == 0 else np.asfortranarray(v.astype(np.float32))
    return g


def F(a):
    return np.asfortranarray(np.array(a, dtype=np.float32))


def Z(rows, cols):
    return np.zeros((rows, cols), dtype=np.float32, order="F")


def geo2d(g):
    """shapes + descriptors of a 2-D golden case."""
    d = GetConvDesc(g["Cin"], g["Cout"], g["ky"], g["kx"], g["sy"], g["sx"], g["py"], g["px"])
    pd = GetConvDesc(g["Cin"], g["Cin"], g["ky"], g["kx"], g["sy"], g["sx"], g["py"], g["px"])
    ish = (g["N"], g["W"], g["H"], g["Cin"])
    fsh = (g["Cout"], g["kx"], g["ky"], g["Cin"])
    tsh = (g["N"], g["modX"], g["modY"], g["Cout"])
    psh = (g["N"], g["modX"], g["modY"], g["Cin"])
    return d, pd, ish, fsh, tsh, psh


def geo3d(g):
    d = GetConvDesc(g["Cin"], g["Cout"], g["ky"], g["kx"], g["sy"], g["sx"], g["py"], g["px"],
                    kernel_size_t=g["kt"], stride_t=g["st"], padding_t=g["pt"])
    pd = GetConvDesc(g["Cin"], g["Cin"], g["ky"], g["kx"], g["sy"], g["sx"], g["py"], g["px"],
                     kernel_size_t=g["kt"], stride_t=g["st"], padding_t=g["pt"])
    ish = (g["N"], g["W"], g["H"], g["Cin"] * g["T"])
    fsh = (g["Cout"], g["kx"], g["ky"], g["Cin"] * g["kt"])
    tsh = (g["N"], g["modX"], g["modY"], g["Cout"] * g["modT"])
    psh = (g["N"], g["modX"], g["modY"], g["Cin"] * g["modT"])
    return d, pd, ish, fsh, tsh, psh
