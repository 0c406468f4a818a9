"""Numpy-backed stand-in for pyiga_b200._device.CudaBackend (TEST INFRASTRUCTURE ONLY).

Pairs with the sequential emulation library built by build_emu.py: buffers are numpy arrays and
"device pointers" are host addresses.  Installed by the `emu` fixture of the CPU tests to check
the host logic of the package and the index logic of the kernels without a GPU.  Nothing in
pyiga_b200 imports this module.
"""
import numpy as np

from pyiga_b200 import _lib


class _Buf(np.ndarray):
    pass


class EmuBackend:
    name = 'emu'

    def __init__(self, libpath):
        self.lib = _lib.bind(libpath)
        self.device_index = 0

    def empty(self, n, dtype=np.float64):
        return np.full(int(n), np.nan if np.dtype(dtype).kind == 'f' else 0, dtype=dtype).view(_Buf)

    def zeros(self, n, dtype=np.float64):
        return np.zeros(int(n), dtype=dtype).view(_Buf)

    def from_host(self, arr, pinned=False):
        return np.array(arr, copy=True).ravel().view(_Buf)

    def to_host(self, buf):
        return np.array(buf, copy=True).view(np.ndarray)

    def ptr(self, buf):
        if buf is None:
            return 0
        if hasattr(buf, 'data_ptr'):        # torch view of a host buffer
            return buf.data_ptr()
        return buf.ctypes.data

    def nbytes(self, buf):
        return buf.nbytes

    def size(self, buf):
        return int(buf.size)

    def itemsize(self, buf):
        return int(buf.itemsize)

    def stream(self):
        return 0

    def synchronize(self):
        pass

    def free_bytes(self):
        return 1 << 34

    def is_buffer(self, x):
        return isinstance(x, _Buf)
