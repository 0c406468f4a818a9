"""Device plumbing: buffers, streams and the native library handle.

PyTorch is used for exactly three things here: allocating device buffers, naming the current
CUDA stream, and host<->device copies.  All numerical work happens in ``libpyiga_b200.so``.
A backend object bundles these services; the package creates :class:`CudaBackend` on first use
and raises if no CUDA device (or no built library) is available.
"""
import ctypes as C

import numpy as np

from . import _lib

_backend = None


class CudaBackend:
    """torch.cuda buffers + the in-tree CUDA library."""
    name = 'cuda'

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('pyiga_b200 needs a CUDA device (B200, sm_100a); '
                               'torch.cuda.is_available() is False and there is no CPU path')
        self.torch = torch
        self.lib = _lib.load()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device('cuda', self.device_index)

    # buffers -----------------------------------------------------------------
    def empty(self, n, dtype=np.float64):
        return self.torch.empty(int(n), dtype=self._tdtype(dtype), device=self.device)

    def zeros(self, n, dtype=np.float64):
        return self.torch.zeros(int(n), dtype=self._tdtype(dtype), device=self.device)

    def _tdtype(self, dtype):
        t = self.torch
        return {np.dtype(np.float64): t.float64, np.dtype(np.int32): t.int32, np.dtype(np.int64): t.int64,
                np.dtype(np.uint64): t.int64, np.dtype(np.uint8): t.uint8}[np.dtype(dtype)]

    def from_host(self, arr, pinned=False):
        arr = np.ascontiguousarray(arr)
        if not arr.flags.writeable:
            arr = arr.copy()
        if arr.dtype == np.uint64:
            arr = arr.view(np.int64)
        t = self.torch.from_numpy(arr)
        if pinned:
            t = t.pin_memory()
        return t.to(self.device, non_blocking=pinned)

    def to_host(self, buf):
        return buf.detach().cpu().numpy()

    def ptr(self, buf):
        return 0 if buf is None else buf.data_ptr()

    def nbytes(self, buf):
        return buf.numel() * buf.element_size()

    def size(self, buf):
        return int(buf.numel())

    def itemsize(self, buf):
        return int(buf.element_size())

    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)

    def free_bytes(self):
        free, _total = self.torch.cuda.mem_get_info(self.device)
        return int(free)

    def is_buffer(self, x):
        return isinstance(x, self.torch.Tensor)


def backend():
    global _backend
    if _backend is None:
        _backend = CudaBackend()
    return _backend


def check(rc):
    _lib.check(backend().lib, rc)


# ---------------------------------------------------------------------------------------------
# spline evaluation on a tensor grid (BSplineFunc / NurbsFunc .grid_eval / .grid_jacobian)
# ---------------------------------------------------------------------------------------------

def eval_spline_on_grid(func, gridaxes, want, keep_on_device=False):
    """Evaluate a spline function or its Jacobian on a tensor grid on the device and return a
    numpy array shaped like the reference's result (``pyiga/bspline.py:874-921``); with
    `keep_on_device` the device buffer itself, viewed in that shape (CUDA backend)."""
    be = backend()
    sdim = func.sdim
    if sdim == 1:
        # a curve: evaluate as a 2D tensor product with a single-point dummy axis
        return _eval_curve(func, gridaxes, want)
    if sdim > 3:
        raise NotImplementedError('grid evaluation is implemented for up to 3 parameters')
    axes = [np.ascontiguousarray(np.squeeze(ax) if np.ndim(ax) != 1 else ax, dtype=np.float64) for ax in gridaxes]
    assert all(ax.ndim == 1 for ax in axes), "Grid axes should be one-dimensional"
    tail = tuple(np.shape(func.coeffs)[sdim:])
    rational = bool(getattr(func, '_rational', False)) or type(func).__name__ == 'NurbsFunc'
    if not rational and (len(tail) > 1 or (len(tail) == 1 and tail[0] > 3)):
        # tensor-valued functions and vectors of more than 3 components: the kernel evaluates up to 3 components
        # per launch — evaluate the flattened components in groups and restore the shape on the host
        from .bspline import BSplineFunc
        flat = np.asarray(func.coeffs, dtype=np.float64).reshape(np.shape(func.coeffs)[:sdim] + (-1,))
        parts = [np.asarray(eval_spline_on_grid(BSplineFunc(func.kvs, flat[..., a:a + 3]), gridaxes, want))
                 for a in range(0, flat.shape[-1], 3)]
        if want == 'value':
            out = np.concatenate(parts, axis=-1)
            return out.reshape(out.shape[:sdim] + tail)
        out = np.concatenate(parts, axis=-2)                # grid + (components, sdim)
        return out.reshape(out.shape[:sdim] + tail + (sdim,))
    desc, keep = _lib.make_geo_desc(func)
    dim = desc.dim
    if dim > 3:
        raise NotImplementedError('rational functions with more than 3 components')
    npts = (C.c_int * sdim)(*[ax.size for ax in axes])
    grids = (_lib.c_double_p * sdim)(*[_lib.as_double_p(ax) for ax in axes])
    total = int(np.prod([ax.size for ax in axes]))
    shape = tuple(ax.size for ax in axes)
    vals = jac = None
    if want == 'value':
        vals = be.empty(total * dim)
    else:
        jac = be.empty(total * dim * sdim)
    check(be.lib.pb200_geo_eval_grid(C.byref(desc), npts, grids, be.ptr(vals), be.ptr(jac),
                                      be.device_index, be.stream()))
    scalar = len(func.output_shape()) == 0
    if keep_on_device and want == 'value':
        out = vals.reshape(shape + (dim,))
        return out[..., 0] if scalar else out
    if want == 'value':
        out = be.to_host(vals).reshape(shape + (dim,))
        return out[..., 0] if scalar else out
    if keep_on_device:
        out = jac.reshape(shape + (dim, sdim))
        return out[..., 0, :] if scalar else out
    out = be.to_host(jac).reshape(shape + (dim, sdim))
    return out[..., 0, :] if scalar else out


def _eval_curve(func, gridaxes, want):
    from . import bspline
    # lift the curve to a surface that is constant in a second parameter
    kv1 = bspline.make_knots(0, 0.0, 1.0, 1)
    lifted = type(func).__new__(type(func))
    lifted.__dict__.update(func.__dict__)
    lifted.kvs = (kv1,) + tuple(func.kvs)
    lifted.sdim = 2
    lifted.coeffs = np.asarray(func.coeffs)[None, ...]
    axes = (np.array([0.5]),) + tuple(gridaxes)
    out = eval_spline_on_grid(lifted, axes, want)[0]
    if want == 'jacobian':
        out = out[..., :1]      # derivative with respect to the curve parameter (last grid axis)
    return out
