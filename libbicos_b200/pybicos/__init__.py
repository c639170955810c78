"""pybicos on the B200-native backend.

Same public surface as the reference's ctypes module (reference pybicos/__init__.py:110-252):
``Config`` (properties nxcorr_threshold / subpixel_step / min_variance / mode / precision /
variant, ``set_no_duplicates()``, ``set_consistency(max_lr_diff, no_dupes)``), the enums
``TransformMode`` / ``Precision`` / ``VariantType``, ``match(stack0, stack1, cfg)`` returning
``(disparity, corrmap)`` numpy arrays, and ``invalid_disparity(dtype)``.

It binds the six symbols of include/pybicos_c.h exported by the ``pybicos_c.so`` next to this
file. That library is ABI-compatible with the reference's, so the reference's own, unmodified
``pybicos/__init__.py`` also works when this ``pybicos_c.so`` is dropped beside it (see
INTEGRATION.md; tests/test_abi.py and tests/test_gpu_parity.py::test_pybicos_drop_in exercise both). One deliberate difference: with the
threshold unset, ``match`` returns ``corrmap=None`` instead of raising on the empty corrmap.
"""

from __future__ import annotations

import ctypes
import enum
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pybicos_c.so")
if not os.path.exists(_SO):
    raise ImportError(f"{_SO} not found: build it with `make -C libbicos_b200/csrc` (no fallback exists)")
_lib = ctypes.CDLL(_SO)


class TransformMode(enum.Enum):
    LIMITED = 0
    FULL = 1


class Precision(enum.Enum):
    SINGLE = 0
    DOUBLE = 1


class VariantType(enum.Enum):
    NO_DUPLICATES = 0
    CONSISTENCY = 1


class _CConfig(ctypes.Structure):
    _fields_ = [(name, ctypes.c_float) for name in ("nxcorr_threshold", "subpixel_step", "min_variance")] + [
        (name, ctypes.c_int) for name in ("mode", "precision", "variant_type", "max_lr_diff", "no_dupes")
    ]


class _CResult(ctypes.Structure):
    _fields_ = [
        ("disparity_data", ctypes.c_void_p), ("disparity_rows", ctypes.c_int),
        ("disparity_cols", ctypes.c_int), ("disparity_type", ctypes.c_int),
        ("corrmap_data", ctypes.c_void_p), ("corrmap_rows", ctypes.c_int),
        ("corrmap_cols", ctypes.c_int), ("corrmap_type", ctypes.c_int),
    ]


_IntP = ctypes.POINTER(ctypes.c_int)
_PtrP = ctypes.POINTER(ctypes.c_void_p)
_lib.BICOS_CreateDefaultConfig.restype = ctypes.POINTER(_CConfig)
_lib.BICOS_FreeConfig.argtypes = [ctypes.POINTER(_CConfig)]
_lib.BICOS_FreeResult.argtypes = [ctypes.POINTER(_CResult)]
_lib.BICOS_Match.restype = ctypes.POINTER(_CResult)
_lib.BICOS_Match.argtypes = [_PtrP, _IntP, _IntP, _IntP, ctypes.c_int] * 2 + [ctypes.POINTER(_CConfig)]
_lib.BICOS_InvalidDisparityFloat.restype = ctypes.c_float
_lib.BICOS_InvalidDisparityInt16.restype = ctypes.c_int16
_lib.BICOS_LastError.restype = ctypes.c_char_p

_DEPTH_OF = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 2}
_DTYPE_OF = {3: np.int16, 5: np.float32, 6: np.float64}


def _optional_float(field):
    def getter(self):
        v = getattr(self._c.contents, field)
        return None if v < 0 else v

    def setter(self, value):
        setattr(self._c.contents, field, -1.0 if value is None else float(value))

    return property(getter, setter)


class Config:
    """Matching parameters; defaults as BICOS::Config (threshold 0.5, LIMITED, SINGLE, NoDuplicates)."""

    def __init__(self):
        self._c = _lib.BICOS_CreateDefaultConfig()
        if not self._c:
            raise MemoryError("BICOS_CreateDefaultConfig failed")

    def __del__(self):
        c, self._c = getattr(self, "_c", None), None
        if c and _lib is not None:  # module globals are already gone at interpreter shutdown
            _lib.BICOS_FreeConfig(c)

    @property
    def _c_config(self):  # name used by the reference module
        return self._c

    nxcorr_threshold = _optional_float("nxcorr_threshold")
    subpixel_step = _optional_float("subpixel_step")
    min_variance = _optional_float("min_variance")

    @property
    def mode(self):
        return TransformMode(self._c.contents.mode)

    @mode.setter
    def mode(self, value):
        self._c.contents.mode = TransformMode(value).value if not isinstance(value, TransformMode) else value.value

    @property
    def precision(self):
        return Precision(self._c.contents.precision)

    @precision.setter
    def precision(self, value):
        self._c.contents.precision = Precision(value).value if not isinstance(value, Precision) else value.value

    @property
    def variant(self):
        c = self._c.contents
        if c.variant_type == VariantType.NO_DUPLICATES.value:
            return "NoDuplicates"
        return {"type": "Consistency", "max_lr_diff": c.max_lr_diff, "no_dupes": bool(c.no_dupes)}

    def set_no_duplicates(self):
        self._c.contents.variant_type = VariantType.NO_DUPLICATES.value

    def set_consistency(self, max_lr_diff=1, no_dupes=False):
        c = self._c.contents
        c.variant_type = VariantType.CONSISTENCY.value
        c.max_lr_diff = int(max_lr_diff)
        c.no_dupes = int(bool(no_dupes))

    def __repr__(self):
        return (f"Config(nxcorr_threshold={self.nxcorr_threshold}, subpixel_step={self.subpixel_step}, "
                f"min_variance={self.min_variance}, mode={self.mode.name}, precision={self.precision.name}, "
                f"variant={self.variant})")


def _marshal(stack):
    keep = [np.ascontiguousarray(img) for img in stack]
    n = len(keep)
    data = (ctypes.c_void_p * n)(*[img.ctypes.data for img in keep])
    rows = (ctypes.c_int * n)(*[img.shape[0] for img in keep])
    cols = (ctypes.c_int * n)(*[img.shape[1] for img in keep])
    try:
        types = (ctypes.c_int * n)(*[_DEPTH_OF[img.dtype] for img in keep])
    except KeyError as e:
        raise ValueError(f"Unsupported numpy dtype: {e.args[0]}") from None
    return keep, (data, rows, cols, types, n)


def _take(ptr, rows, cols, type_code):
    if not ptr or rows * cols == 0:
        return None
    dtype = np.dtype(_DTYPE_OF[type_code & 7])
    buf = (ctypes.c_byte * (rows * cols * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(rows, cols).copy()


def match(stack0, stack1, cfg=None):
    """Match two lists of equally sized uint8/uint16 images; returns (disparity, corrmap)."""
    if not len(stack0) or not len(stack1):
        raise ValueError("Empty image stacks")
    cfg = cfg or Config()
    keep0, args0 = _marshal(stack0)
    keep1, args1 = _marshal(stack1)
    res = _lib.BICOS_Match(*args0, *args1, cfg._c)
    if not res:
        raise RuntimeError("BICOS matching failed: " + _lib.BICOS_LastError().decode())
    try:
        r = res.contents
        disparity = _take(r.disparity_data, r.disparity_rows, r.disparity_cols, r.disparity_type)
        corrmap = _take(r.corrmap_data, r.corrmap_rows, r.corrmap_cols, r.corrmap_type)
    finally:
        _lib.BICOS_FreeResult(res)
    del keep0, keep1
    return disparity, corrmap


def invalid_disparity(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return _lib.BICOS_InvalidDisparityFloat()
    if dtype == np.int16:
        return _lib.BICOS_InvalidDisparityInt16()
    raise ValueError(f"Unsupported dtype for invalid_disparity: {dtype}")
