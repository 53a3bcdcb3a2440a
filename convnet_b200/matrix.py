"""CUDAMatrix — host-side handle of a column-major fp32 device matrix (mirror of
cudamat.CUDAMatrix, cudamat/cudamat.py:187-330, reduced to what the conv path needs).

Device memory is a torch CUDA tensor (PyTorch is plumbing here: allocation, streams,
torch.distributed).  `p_mat` / `p_shape4d` are what the C ABI takes.
"""
import ctypes as ct

import numpy as np
import torch

from .abi import Shape4D, cudamat


class CUDAMatrix:
    def __init__(self, rows, cols, shape4d=None, device="cuda:0", storage=None):
        self.rows, self.cols = int(rows), int(cols)
        n = self.rows * self.cols
        if storage is None:
            storage = torch.zeros(max(n, 1), dtype=torch.float32, device=device)
        assert storage.dtype == torch.float32 and storage.is_contiguous() and storage.numel() >= n
        self.storage = storage           # 1-D tensor; element (r, c) at storage[r + rows*c]
        self.mat = cudamat()
        self.mat.data_host = None
        self.mat.data_device = storage.data_ptr()
        self.mat.on_device, self.mat.on_host = 1, 0
        self.mat.size[0], self.mat.size[1] = self.rows, self.cols
        self.mat.is_trans, self.mat.owns_data, self.mat.tex_obj = 0, 0, 0
        self.p_mat = ct.pointer(self.mat)
        self.set_shape4d(shape4d if shape4d is not None else (self.rows, self.cols, 1, 1))

    # -- shape ------------------------------------------------------------------------------
    def set_shape4d(self, s):
        self.shape4d = tuple(int(v) for v in s)
        self._s4 = Shape4D.of(*self.shape4d)
        self.p_shape4d = ct.pointer(self._s4)

    @property
    def shape(self):
        return (self.rows, self.cols)

    @property
    def ptr(self):
        return self.storage.data_ptr()

    # -- host <-> device ----------------------------------------------------------------------
    @classmethod
    def from_numpy(cls, a, shape4d=None, device="cuda:0"):
        a = np.asarray(a, dtype=np.float32)
        assert a.ndim == 2
        flat = np.ascontiguousarray(a.T).reshape(-1)          # column-major linearisation
        t = torch.from_numpy(flat).to(device)
        return cls(a.shape[0], a.shape[1], shape4d, device, t)

    def asarray(self):
        flat = self.storage[: self.rows * self.cols].detach().cpu().numpy()
        return np.asfortranarray(flat.reshape(self.cols, self.rows).T)

    def tensor2d(self):
        """[cols, rows] row-major torch view == column-major [rows, cols]."""
        return self.storage[: self.rows * self.cols].view(self.cols, self.rows)

    def fill_(self, v):
        self.storage.fill_(v)
        return self
