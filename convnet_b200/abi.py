"""ctypes mirror of the C-ABI structs (include/cudamat_abi.h).

Same layouts as the reference's ctypes view, cudamat/cudamat.py:127-157
(`cudamat`, `Shape4D`, `ConvDesc`), and the same `GetConvDesc` helper signature
(cudamat/cudamat.py:159-185): user-facing padding is POSITIVE and is negated
into the descriptor, exactly like src/edge.cc:97-99.
"""
import ctypes as ct


class cudamat(ct.Structure):
    _fields_ = [
        ("data_host", ct.POINTER(ct.c_float)),
        ("data_device", ct.c_void_p),
        ("on_device", ct.c_int),
        ("on_host", ct.c_int),
        ("size", ct.c_int * 2),
        ("is_trans", ct.c_int),
        ("owns_data", ct.c_int),
        ("tex_obj", ct.c_ulonglong),
    ]


class Shape4D(ct.Structure):
    _fields_ = [("shape", ct.c_int * 4)]

    @classmethod
    def of(cls, a, b, c, d):
        s = cls()
        s.shape[0], s.shape[1], s.shape[2], s.shape[3] = int(a), int(b), int(c), int(d)
        return s

    def tuple(self):
        return tuple(self.shape[i] for i in range(4))


class ConvDesc(ct.Structure):
    _fields_ = [
        ("num_input_channels", ct.c_int),
        ("num_output_channels", ct.c_int),
        ("kernel_size_y", ct.c_int),
        ("kernel_size_x", ct.c_int),
        ("kernel_size_t", ct.c_int),
        ("stride_y", ct.c_int),
        ("stride_x", ct.c_int),
        ("stride_t", ct.c_int),
        ("padding_y", ct.c_int),
        ("padding_x", ct.c_int),
        ("padding_t", ct.c_int),
        ("input_channel_begin", ct.c_int),
        ("input_channel_end", ct.c_int),
        ("output_channel_begin", ct.c_int),
        ("output_channel_end", ct.c_int),
        ("num_groups", ct.c_int),
    ]

    def copy(self):
        c = ConvDesc()
        ct.memmove(ct.byref(c), ct.byref(self), ct.sizeof(ConvDesc))
        return c


assert ct.sizeof(cudamat) == 48 and ct.sizeof(Shape4D) == 16 and ct.sizeof(ConvDesc) == 64


def GetConvDesc(num_input_channels, num_output_channels, kernel_size_y, kernel_size_x,
                stride_y, stride_x, padding_y, padding_x,
                kernel_size_t=1, stride_t=1, padding_t=0,
                input_channel_begin=0, input_channel_end=0,
                output_channel_begin=0, output_channel_end=0, num_groups=1):
    """cudamat/cudamat.py:159-185: positive paddings in, negated into the descriptor."""
    d = ConvDesc()
    d.num_input_channels = num_input_channels
    d.num_output_channels = num_output_channels
    d.kernel_size_y, d.kernel_size_x, d.kernel_size_t = kernel_size_y, kernel_size_x, kernel_size_t
    d.stride_y, d.stride_x, d.stride_t = stride_y, stride_x, stride_t
    d.padding_y, d.padding_x, d.padding_t = -padding_y, -padding_x, -padding_t
    d.input_channel_begin = input_channel_begin
    d.input_channel_end = input_channel_end if input_channel_end else num_input_channels
    d.output_channel_begin = output_channel_begin
    d.output_channel_end = output_channel_end if output_channel_end else num_output_channels
    d.num_groups = num_groups
    return d


def num_modules(image_size, kernel_size, stride, padding_pos):
    """src/edge.cc:108-114."""
    return (image_size + 2 * padding_pos - kernel_size) // stride + 1
