"""Zero-copy device pointers out of DLPack capsules (ctypes only, no framework import).

Anything that speaks DLPack works as an argument of the ops in this package: torch tensors, CuPy
arrays, or TensorFlow tensors via `tf.experimental.dlpack.to_dlpack(t)` (a raw capsule).  The capsule
is consumed the DLPack way (renamed to "used_dltensor") and its deleter runs when the returned
`DeviceArray` is released, so the producer's memory stays alive exactly as long as we hold it.
"""
import ctypes

import numpy as np

kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13
_CODES = {0: 'int', 1: 'uint', 2: 'float', 4: 'bfloat'}


class DLDevice(ctypes.Structure):
    _fields_ = [('device_type', ctypes.c_int), ('device_id', ctypes.c_int)]


class DLDataType(ctypes.Structure):
    _fields_ = [('code', ctypes.c_uint8), ('bits', ctypes.c_uint8), ('lanes', ctypes.c_uint16)]


class DLTensor(ctypes.Structure):
    _fields_ = [('data', ctypes.c_void_p), ('device', DLDevice), ('ndim', ctypes.c_int), ('dtype', DLDataType),
                ('shape', ctypes.POINTER(ctypes.c_int64)), ('strides', ctypes.POINTER(ctypes.c_int64)),
                ('byte_offset', ctypes.c_uint64)]


class DLManagedTensor(ctypes.Structure):
    pass


_DELETER = ctypes.CFUNCTYPE(None, ctypes.POINTER(DLManagedTensor))
DLManagedTensor._fields_ = [('dl_tensor', DLTensor), ('manager_ctx', ctypes.c_void_p), ('deleter', _DELETER)]

_api = ctypes.pythonapi
_api.PyCapsule_IsValid.restype = ctypes.c_int
_api.PyCapsule_IsValid.argtypes = [ctypes.py_object, ctypes.c_char_p]
_api.PyCapsule_GetPointer.restype = ctypes.c_void_p
_api.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
_api.PyCapsule_SetName.restype = ctypes.c_int
_api.PyCapsule_SetName.argtypes = [ctypes.py_object, ctypes.c_char_p]


class DeviceArray:
    """A borrowed, contiguous device buffer: pointer + shape + dtype string ('float32', 'uint8', ...)."""

    __slots__ = ('ptr', 'shape', 'dtype', 'itemsize', 'device_type', 'device_id', '_managed', '_capsule', '_owner')

    def __init__(self, ptr, shape, dtype, itemsize, device_type, device_id, managed=None, capsule=None, owner=None):
        self.ptr, self.shape, self.dtype, self.itemsize = ptr, tuple(shape), dtype, itemsize
        self.device_type, self.device_id = device_type, device_id
        self._managed, self._capsule, self._owner = managed, capsule, owner

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def nbytes(self):
        return self.size * self.itemsize

    def release(self):
        managed, self._managed = self._managed, None
        if managed is not None and managed.contents.deleter:
            managed.contents.deleter(managed)
        self._capsule = self._owner = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def from_capsule(capsule, owner=None):
    if not _api.PyCapsule_IsValid(capsule, b'dltensor'):
        raise TypeError('expected an unconsumed DLPack capsule named "dltensor"')
    raw = _api.PyCapsule_GetPointer(capsule, b'dltensor')
    managed = ctypes.cast(raw, ctypes.POINTER(DLManagedTensor))
    t = managed.contents.dl_tensor
    _api.PyCapsule_SetName(capsule, b'used_dltensor')       # we own the deleter call from here on
    shape = [t.shape[i] for i in range(t.ndim)]
    if t.dtype.lanes != 1 or t.dtype.code not in _CODES:
        raise TypeError(f'unsupported DLPack dtype code={t.dtype.code} lanes={t.dtype.lanes}')
    if t.strides:
        expect = 1
        for i in range(t.ndim - 1, -1, -1):
            if shape[i] != 1 and t.strides[i] != expect:
                arr = DeviceArray(0, shape, '', 0, 0, 0, managed)
                arr.release()
                raise ValueError('DLPack tensor must be C-contiguous')
            expect *= shape[i]
    dtype = f'{_CODES[t.dtype.code]}{t.dtype.bits}'
    ptr = (t.data or 0) + t.byte_offset
    return DeviceArray(ptr, shape, dtype, t.dtype.bits // 8, t.device.device_type, t.device.device_id, managed, capsule, owner)


_TORCH_DTYPES = None


def _from_torch(t):
    """Same DeviceArray a torch tensor's capsule would give, without building the capsule (the hot loop
    calls this ~10x per launch; the capsule round trip costs ~4 us each)."""
    global _TORCH_DTYPES
    import torch
    if _TORCH_DTYPES is None:
        _TORCH_DTYPES = {torch.float32: ('float32', 4), torch.float64: ('float64', 8), torch.int32: ('int32', 4),
                         torch.int64: ('int64', 8), torch.uint8: ('uint8', 1), torch.int8: ('int8', 1),
                         torch.float16: ('float16', 2), torch.bfloat16: ('bfloat16', 2), torch.int16: ('int16', 2)}
    if t.dtype not in _TORCH_DTYPES:
        raise TypeError(f'unsupported torch dtype {t.dtype}')
    if not t.is_contiguous():
        raise ValueError('DLPack tensor must be C-contiguous')
    name, size = _TORCH_DTYPES[t.dtype]
    dev_type = kDLCUDA if t.is_cuda else kDLCPU
    return DeviceArray(t.data_ptr(), t.shape, name, size, dev_type, t.device.index or 0, owner=t)


def as_device_array(obj, dtype=None, allow_host=False):
    """Borrow `obj` (a DLPack capsule or any object with __dlpack__) as a DeviceArray."""
    if isinstance(obj, DeviceArray):
        arr = obj
    elif type(obj).__module__.split('.')[0] == 'torch' and hasattr(obj, 'data_ptr') and hasattr(obj, 'is_cuda'):
        arr = _from_torch(obj.detach() if obj.requires_grad else obj)
    elif type(obj).__name__ == 'PyCapsule':
        arr = from_capsule(obj)
    elif hasattr(obj, '__dlpack__'):
        arr = from_capsule(obj.__dlpack__(), owner=obj)
    else:
        raise TypeError(f'{type(obj).__name__} does not speak DLPack (need __dlpack__ or a "dltensor" capsule)')
    if not allow_host and arr.device_type not in (kDLCUDA, kDLCUDAManaged):
        raise ValueError('expected a CUDA tensor: this package has no CPU path '
                         f'(got DLPack device_type={arr.device_type})')
    if dtype is not None and arr.dtype != dtype:
        raise TypeError(f'expected dtype {dtype}, got {arr.dtype}')
    return arr


def host_numpy(arr):
    """View a HOST DeviceArray (kDLCPU / kDLCUDAHost) as numpy, for tests of the capsule reader."""
    if arr.device_type not in (kDLCPU, kDLCUDAHost):
        raise ValueError('not a host tensor')
    buf = (ctypes.c_char * arr.nbytes).from_address(arr.ptr)
    return np.frombuffer(buf, dtype=np.dtype(arr.dtype)).reshape(arr.shape)
