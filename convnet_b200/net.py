"""ctypes door onto the native host code (convnet_b200/host: Matrix / Edge / ConvNet / GradChecker /
DataParallelSync -> lib/libconvnet_b200_host.so).  Python only launches; sequencing, memory and the
NCCL gradient sync live in C++ like the reference's src/convnet.cc."""
import ctypes as ct
import os

from . import lib as _lib

HOST_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libconvnet_b200_host.so")
_host = None


def load_host():
    global _host
    if _host is None:
        _lib.load()                    # the kernel library first (RTLD_GLOBAL not needed: host lib has an rpath)
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError("convnet_b200: %s is missing - run __graft_entry__.build()" % HOST_LIB_PATH)
        H = ct.CDLL(HOST_LIB_PATH)
        vp, i, ll, d, f = ct.c_void_p, ct.c_int, ct.c_longlong, ct.c_double, ct.c_float
        sig = {
            "cnb_net_create": ([ct.c_char_p, i, ct.c_uint, i], vp), "cnb_net_destroy": ([vp], None),
            "cnb_net_num_params": ([vp], ll), "cnb_net_num_edges": ([vp], i), "cnb_net_edge_name": ([vp, i], ct.c_char_p),
            "cnb_net_edge_flops": ([vp, i], d), "cnb_net_edge_offset": ([vp, i], ll), "cnb_net_edge_size": ([vp, i], ll),
            "cnb_net_flops_fprop": ([vp], d), "cnb_net_flops_train": ([vp], d),
            "cnb_net_input": ([vp], vp), "cnb_net_input_floats": ([vp], ll), "cnb_net_labels": ([vp], vp),
            "cnb_net_output": ([vp], vp), "cnb_net_num_classes": ([vp], i), "cnb_net_params": ([vp], vp),
            "cnb_net_grads": ([vp], vp), "cnb_net_layer_state": ([vp, i], vp), "cnb_net_layer_floats": ([vp, i], ll),
            "cnb_net_num_layers": ([vp], i), "cnb_net_device_loss": ([vp], vp),
            "cnb_net_fprop": ([vp, i], None), "cnb_net_bprop": ([vp], None), "cnb_net_update": ([vp], None),
            "cnb_net_loss": ([vp], f), "cnb_net_train_step": ([vp, ct.POINTER(f)], None),
            "cnb_net_trace_step": ([vp, ct.POINTER(f), i], i),
            "cnb_data_create": ([i, i, i, i, i, i, i, i, ct.c_ulonglong], vp), "cnb_data_destroy": ([vp], None),
            "cnb_data_upload": ([vp, vp, i, i], None), "cnb_data_get_batch": ([vp, vp, i, i], None),
            "cnb_data_last_noise": ([vp, ct.POINTER(f), i], i),
            "cnb_data_view_offset": ([i, i, i, ct.POINTER(i), ct.POINTER(i)], None),
            "cnb_dp_unique_id": ([ct.c_char_p], i), "cnb_net_dp_init": ([vp, i, i, ct.c_char_p, ll], i),
            "cnb_plan_buckets": ([i, ct.POINTER(ll), ct.POINTER(ll), ll, i, ct.POINTER(ll), ct.POINTER(ll), ct.POINTER(i)], i),
            "cnb_model_edge_params": ([ct.c_char_p, i, i, ct.POINTER(ll)], i),
            "cnb_net_grad_check": ([vp, ct.c_uint, i, ct.c_char_p, ct.POINTER(f), ct.POINTER(f), ct.POINTER(f)], i),
        }
        for name, (args, res) in sig.items():
            fn = getattr(H, name)
            fn.argtypes, fn.restype = args, res
        _host = H
    return _host


class Net:
    """A chain ConvNet built natively ("alexnet" | "lenet" | "c3d" | "tiny")."""

    def __init__(self, model, batch_size, seed=42, grad_checker=False):
        self.H = load_host()
        self.h = self.H.cnb_net_create(model.encode(), batch_size, seed, int(grad_checker))
        self.batch_size = batch_size
        self.model = model

    def close(self):
        if self.h:
            self.H.cnb_net_destroy(self.h)
            self.h = None

    # --- sizes
    num_params = property(lambda s: s.H.cnb_net_num_params(s.h))
    num_classes = property(lambda s: s.H.cnb_net_num_classes(s.h))
    input_floats = property(lambda s: s.H.cnb_net_input_floats(s.h))
    flops_fprop = property(lambda s: s.H.cnb_net_flops_fprop(s.h))
    flops_train = property(lambda s: s.H.cnb_net_flops_train(s.h))

    def edges(self):
        return [(self.H.cnb_net_edge_name(self.h, i).decode(), self.H.cnb_net_edge_flops(self.h, i),
                 self.H.cnb_net_edge_offset(self.h, i), self.H.cnb_net_edge_size(self.h, i))
                for i in range(self.H.cnb_net_num_edges(self.h))]

    # --- device buffers as torch tensors (zero-copy views)
    def _view(self, ptr, n, dtype):
        import torch
        # torch has no from_address; go through the CUDA array interface

        class _W:
            pass
        w = _W()
        w.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4" if dtype == "f" else "<i4",
                                      "data": (ptr, False), "version": 2}
        return torch.as_tensor(w, device="cuda")

    def input_tensor(self):
        return self._view(self.H.cnb_net_input(self.h), self.input_floats, "f")

    def labels_tensor(self):
        return self._view(self.H.cnb_net_labels(self.h), self.batch_size, "i")

    def output_tensor(self):
        return self._view(self.H.cnb_net_output(self.h), self.batch_size * self.num_classes, "f")

    def params_tensor(self):
        return self._view(self.H.cnb_net_params(self.h), self.num_params, "f")

    def grads_tensor(self):
        return self._view(self.H.cnb_net_grads(self.h), self.num_params, "f")

    def layer_state(self, i):
        return self._view(self.H.cnb_net_layer_state(self.h, i), self.H.cnb_net_layer_floats(self.h, i), "f")

    # --- compute
    def fprop(self, train=False):
        self.H.cnb_net_fprop(self.h, int(train))

    def bprop(self):
        self.H.cnb_net_bprop(self.h)

    def update(self):
        self.H.cnb_net_update(self.h)

    def loss(self):
        return self.H.cnb_net_loss(self.h)

    def train_step(self, want_loss=True):
        if want_loss:
            v = ct.c_float(0)
            self.H.cnb_net_train_step(self.h, ct.byref(v))
            return v.value
        self.H.cnb_net_train_step(self.h, None)
        return None

    def trace_step(self):
        """One extra training step with timing events on the three streams (ConvNet::TraceStep): device milliseconds since
        the step began of the pipeline's milestones, and per gradient bucket its size, exchange window and SGD end."""
        buf = (ct.c_float * 512)()
        n = min(512, self.H.cnb_net_trace_step(self.h, buf, 512))
        v = [buf[k] for k in range(n)]
        out = {"fprop_end_ms": v[0], "bprop_compute_end_ms": v[1], "step_end_ms": v[2], "buckets": []}
        for b in range(int(v[3])):
            mb, c0, c1, s1 = v[4 + 4 * b:8 + 4 * b]
            out["buckets"].append({"MB": round(mb, 3), "exchange_begin_ms": c0, "exchange_end_ms": c1, "sgd_end_ms": s1})
        return out

    def dp_init(self, rank, world, id_bytes, bucket_floats=32 << 20):
        assert len(id_bytes) == 128
        rc = self.H.cnb_net_dp_init(self.h, rank, world, bytes(id_bytes), bucket_floats)
        if rc != 0:
            raise RuntimeError("NCCL initialisation failed (libnccl.so.2 not loadable?)")

    def grad_check(self, seed=1, cap=64):
        names = ct.create_string_buffer(64 * cap)
        eps, dw, db = (ct.c_float * cap)(), (ct.c_float * cap)(), (ct.c_float * cap)()
        n = self.H.cnb_net_grad_check(self.h, seed, cap, names, eps, dw, db)
        out = []
        for k in range(n):
            out.append((names.raw[64 * k:64 * (k + 1)].split(b"\0")[0].decode(), eps[k], dw[k], db[k]))
        return out


def model_edge_params(model, batch=1):
    """parameter count of every edge of a model (host-only: no device memory is touched)."""
    H = load_host()
    buf = (ct.c_longlong * 64)()
    n = H.cnb_model_edge_params(model.encode(), batch, 64, buf)
    return [buf[k] for k in range(n)]


def plan_buckets(edge_sizes, bucket_floats):
    """(lo, hi, trigger_edge) gradient buckets over the flat 128-float-padded parameter buffer (convnet.cc PlanBuckets)."""
    H = load_host()
    n = len(edge_sizes)
    offs, total = [], 0
    for sz in edge_sizes:
        offs.append(total)
        total += (sz + 127) // 128 * 128
    A = ct.c_longlong * n
    cap = n + 1
    lo, hi, trig = (ct.c_longlong * cap)(), (ct.c_longlong * cap)(), (ct.c_int * cap)()
    k = H.cnb_plan_buckets(n, A(*offs), A(*edge_sizes), bucket_floats, cap, lo, hi, trig)
    return [(lo[j], hi[j], trig[j]) for j in range(k)], total


def dp_unique_id():
    buf = ct.create_string_buffer(128)
    if load_host().cnb_dp_unique_id(buf) != 0:
        raise RuntimeError("ncclGetUniqueId failed")
    return buf.raw


class DataIterator:
    """Device side of the reference's input pipeline (host/data.h; src/datahandler.cc:146-200, 520-568): a chunk of images
    resident on the GPU, and per minibatch a random (or centre / corner) crop + mirror into the net's input layer."""

    def __init__(self, chunk_size, channels, image_size, gpu_image_size, translate=True, flip=True, seed=1):
        self.H = load_host()
        isy, isx = (image_size, image_size) if isinstance(image_size, int) else image_size
        gy, gx = (gpu_image_size, gpu_image_size) if isinstance(gpu_image_size, int) else gpu_image_size
        self.chunk_size, self.dims = chunk_size, channels * isy * isx
        self.h = self.H.cnb_data_create(chunk_size, channels, isy, isx, gy, gx, int(translate), int(flip), seed)

    def close(self):
        if self.h:
            self.H.cnb_data_destroy(self.h)
            self.h = None

    def upload(self, host_tensor, first=0):
        """host_tensor: float32 CPU tensor [count, channels, rows, cols] (pinned for an asynchronous copy)"""
        import torch
        t = host_tensor.contiguous()
        assert t.dtype == torch.float32 and t.device.type == "cpu" and t[0].numel() == self.dims
        self.H.cnb_data_upload(self.h, t.data_ptr(), first, t.shape[0])
        self._keep = t                                   # the copy is asynchronous

    def get_batch(self, net, start=0, multiplicity_id=0):
        self.H.cnb_data_get_batch(self.h, net.h, start, multiplicity_id)

    def last_noise(self, batch):
        buf = (ct.c_float * (3 * batch))()
        n = self.H.cnb_data_last_noise(self.h, buf, 3 * batch)
        assert n == batch
        v = [buf[k] for k in range(3 * batch)]
        return v[:batch], v[batch:2 * batch], v[2 * batch:]


def view_offset(multiplicity_id, max_offset_x, max_offset_y):
    """(w, h) of the deterministic crop number multiplicity_id (host logic only, no GPU)"""
    H = load_host()
    w, h = ct.c_int(0), ct.c_int(0)
    H.cnb_data_view_offset(multiplicity_id, max_offset_x, max_offset_y, ct.byref(w), ct.byref(h))
    return w.value, h.value
