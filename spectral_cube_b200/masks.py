"""
Mask classes with the reference's names and semantics (spectral_cube/masks.py:101-803), held
as a small expression tree that is lowered to the C ABI's ``sc_mask_desc`` and evaluated in
registers inside the CUDA kernels (no boolean cube is materialised for lazy masks).

Data-carrying masks keep references to **device tensors** (torch, used only as memory
holders).  ``include()`` / ``exclude()`` materialise the predicate on the GPU and return a
numpy bool array, like the reference.
"""
import operator

import numpy as np

from . import _lib

_OPS = {operator.gt: _lib.GT, operator.ge: _lib.GE, operator.lt: _lib.LT,
        operator.le: _lib.LE, operator.eq: _lib.EQ, operator.ne: _lib.NE}


def _torch():
    return _lib.require_cuda()


def _as_device(arr, dtype=None):
    torch = _torch()
    if isinstance(arr, torch.Tensor):
        t = arr
    else:
        t = torch.from_numpy(np.ascontiguousarray(arr))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.cuda()


def is_broadcastable_and_smaller(shp1, shp2):
    """masks.py:31-45"""
    if len(shp1) < len(shp2):
        return False
    for a, b in zip(shp1[::-1], shp2[::-1]):
        if a != 1 and b != 1 and a != b:
            return False
        if a < b:
            return False
    return True


class _Lowering(object):
    """Accumulates nodes children-first while walking a mask tree."""

    def __init__(self, cube_tensor):
        self.cube = cube_tensor
        self.nodes = []
        self.keep = []          # tensors that must outlive the call

    def add(self, **kw):
        if len(self.nodes) >= _lib.SC_MASK_MAX_NODES:
            raise ValueError("mask expression has more than %d nodes; materialise part of it "
                             "with BooleanArrayMask(mask.include(), wcs)" % _lib.SC_MASK_MAX_NODES)
        self.nodes.append(kw)
        return len(self.nodes) - 1

    def data_fields(self, data):
        """(ptr, ds_c, ds_y) for a float32 data tensor the mask was built on."""
        if data is None or (data.data_ptr() == self.cube.data_ptr() and data.stride() == self.cube.stride()
                            and data.device == self.cube.device):
            return dict(data=None, ds_c=0, ds_y=0)
        if tuple(data.shape) != tuple(self.cube.shape):
            raise ValueError("mask data shape %s does not match cube shape %s"
                             % (tuple(data.shape), tuple(self.cube.shape)))
        if data.device != self.cube.device:
            data = data.to(self.cube.device)             # a mask built on another GPU's copy of the data
        if data.stride(2) != 1:
            data = data.contiguous()
        self.keep.append(data)
        return dict(data=data.data_ptr(), ds_c=data.stride(0), ds_y=data.stride(1))

    def array_fields(self, arr):
        """(ptr, strides with 0 on broadcast axes) for an array broadcastable to the cube."""
        t = arr if arr.device == self.cube.device else arr.to(self.cube.device)
        while t.dim() < 3:
            t = t.unsqueeze(0)
        e = t.expand(tuple(self.cube.shape))
        self.keep.append(t)
        return dict(array=t.data_ptr(), as_c=e.stride(0), as_y=e.stride(1), as_x=e.stride(2))

    def descriptor(self):
        d = _lib.MaskDesc()
        d.n_nodes = len(self.nodes)
        for i, kw in enumerate(self.nodes):
            n = d.nodes[i]
            for k, v in kw.items():
                setattr(n, k, v)
        return d


def lower_mask(mask, cube_tensor):
    """mask (MaskBase or None) -> (MaskDesc, keepalive list)."""
    low = _Lowering(cube_tensor)
    if mask is not None:
        mask._lower(low)
    return low.descriptor(), low.keep


class MaskBase(object):
    """masks.py:101-334"""
    shape = None

    # -- evaluation (on the GPU) -----------------------------------------------------------
    def _include_tensor(self, data=None):
        """uint8 device tensor of the include predicate over the whole mask shape."""
        torch = _torch()
        lib = _lib.load()
        ref = data if data is not None else self._reference_tensor()
        if ref is None:
            raise ValueError("this mask needs `data` to know its shape")
        ref = _as_device(ref, torch.float32)
        if ref.stride(2) != 1:
            ref = ref.contiguous()
        desc, keep = lower_mask(self, ref)
        out = torch.empty(tuple(ref.shape), dtype=torch.uint8, device=ref.device)
        nchan, ny, nx = ref.shape
        with _lib.on_device_of(ref):
            _lib.check(lib.sc_mask_include(ref.data_ptr(), nchan, ny, nx, ref.stride(0), ref.stride(1),
                                           desc, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return out

    def _reference_tensor(self):
        return None

    def include(self, data=None, wcs=None, view=()):
        return self._include_tensor(data).bool().cpu().numpy()[view]

    def exclude(self, data=None, wcs=None, view=()):
        return np.logical_not(self.include(data=data, wcs=wcs, view=view))

    def any(self):
        return bool(self._include_tensor().any().item())

    def _filled(self, data, wcs=None, fill=np.nan, view=(), **kwargs):
        """masks.py:197-237: data with excluded voxels replaced by ``fill`` (numpy array)."""
        torch = _torch()
        d = _as_device(data, torch.float32)
        inc = self._include_tensor(d).bool()
        out = torch.where(inc, d, torch.full((), float(fill), dtype=d.dtype, device=d.device))
        return out.cpu().numpy()[view]

    def _flattened(self, data, wcs=None, view=()):
        torch = _torch()
        d = _as_device(data, torch.float32)
        inc = self._include_tensor(d).bool()
        if view != ():
            return d.cpu().numpy()[view][inc.cpu().numpy()[view]]
        return d[inc].cpu().numpy()

    # -- operators (masks.py:239-249) ------------------------------------------------------------
    def __and__(self, other):
        return CompositeMask(self, other, operation='and')

    def __or__(self, other):
        return CompositeMask(self, other, operation='or')

    def __xor__(self, other):
        return CompositeMask(self, other, operation='xor')

    def __invert__(self):
        return InvertedMask(self)

    def _validate_wcs(self, new_data=None, new_wcs=None, **kwargs):
        if new_data is not None and self.shape is not None:
            if not is_broadcastable_and_smaller(tuple(new_data.shape), tuple(self.shape)):
                raise ValueError("data shape cannot be broadcast to match mask shape")


def _slice3(view):
    """Normalise a cube view to three slices (masks.py `__getitem__` / wcs_utils.slice_wcs take the same)."""
    if not isinstance(view, tuple):
        view = (view,)
    if any(v is Ellipsis for v in view) or len(view) > 3 or not all(isinstance(v, slice) for v in view):
        raise NotImplementedError("only views made of up to three slices are supported")
    return tuple(view) + (slice(None),) * (3 - len(view))


def _slice_broadcast(t, view):
    """Slice a tensor that is broadcastable to the cube: axes of length 1 stay as they are."""
    view = _slice3(view)
    if t.dim() < 3:
        view = view[3 - t.dim():]
    return t[tuple(slice(None) if t.shape[i] == 1 else v for i, v in enumerate(view))]


class InvertedMask(MaskBase):
    """masks.py:337-361"""

    def __init__(self, mask):
        self._mask = mask

    @property
    def shape(self):
        return self._mask.shape

    def _reference_tensor(self):
        return self._mask._reference_tensor()

    def __getitem__(self, view):
        return InvertedMask(self._mask[view])

    def _lower(self, low):
        a = self._mask._lower(low)
        return low.add(kind=_lib.MASK_NOT, a=a)


class CompositeMask(MaskBase):
    """masks.py:364-454"""

    def __init__(self, mask1, mask2, operation='and'):
        if isinstance(mask1, np.ndarray) and isinstance(mask2, MaskBase) and hasattr(mask2, 'shape'):
            mask1 = BooleanArrayMask(mask1, wcs=None, shape=mask2.shape)
        elif isinstance(mask2, np.ndarray) and isinstance(mask1, MaskBase) and hasattr(mask1, 'shape'):
            mask2 = BooleanArrayMask(mask2, wcs=None, shape=mask1.shape)
        if operation not in ('and', 'or', 'xor'):
            raise ValueError("Operation '{0}' not supported".format(operation))
        if (mask1.shape is not None and mask2.shape is not None and
                not is_broadcastable_and_smaller(tuple(mask1.shape), tuple(mask2.shape)) and
                not is_broadcastable_and_smaller(tuple(mask2.shape), tuple(mask1.shape))):
            raise ValueError("The two masks cannot be broadcast to the same shape")
        self._mask1, self._mask2, self._operation = mask1, mask2, operation

    @property
    def shape(self):
        a, b = self._mask1.shape, self._mask2.shape
        if a is None:
            return b
        if b is None:
            return a
        return tuple(max(i, j) for i, j in zip(a, b))

    def _reference_tensor(self):
        r = self._mask1._reference_tensor()
        return r if r is not None else self._mask2._reference_tensor()

    def __getitem__(self, view):
        return CompositeMask(self._mask1[view], self._mask2[view], operation=self._operation)

    def _lower(self, low):
        a = self._mask1._lower(low)
        b = self._mask2._lower(low)
        kind = {'and': _lib.MASK_AND, 'or': _lib.MASK_OR, 'xor': _lib.MASK_XOR}[self._operation]
        return low.add(kind=kind, a=a, b=b)


class BooleanArrayMask(MaskBase):
    """masks.py:457-584.  ``mask`` may be smaller than the cube but broadcastable to it."""

    def __init__(self, mask, wcs=None, include=True, shape=None):
        torch = _torch()
        if isinstance(mask, torch.Tensor):
            t = mask.to(torch.uint8) if mask.dtype != torch.uint8 else mask
        else:
            m = np.asarray(mask)
            if m.dtype != bool:
                raise TypeError("BooleanArrayMask requires a boolean array")
            t = torch.from_numpy(np.ascontiguousarray(m).view(np.uint8))
        self._mask = t.cuda()
        self._mask_type = 'include' if include else 'exclude'
        self._wcs = wcs
        if shape is not None:
            if not is_broadcastable_and_smaller(tuple(shape), tuple(self._mask.shape)):
                raise ValueError("Mask cannot be broadcast to the specified shape.")
        self._shape = tuple(shape) if shape is not None else tuple(self._mask.shape)

    @property
    def shape(self):
        return self._shape

    def _reference_tensor(self):
        torch = _torch()
        return torch.zeros(self._shape, dtype=torch.float32, device=self._mask.device)

    def __getitem__(self, view):
        """masks.py:559-566: the sliced array (a view of the same device memory)."""
        sub = _slice_broadcast(self._mask, view)
        new = BooleanArrayMask.__new__(BooleanArrayMask)
        new._mask, new._mask_type, new._wcs = sub, self._mask_type, self._wcs
        full = _torch().empty(self._shape, dtype=_torch().uint8, device='meta')[_slice3(view)]
        new._shape = tuple(full.shape)
        return new

    def _lower(self, low):
        i = low.add(kind=_lib.MASK_BOOL, array_dtype=_lib.U8, **low.array_fields(self._mask))
        if self._mask_type == 'exclude':
            i = low.add(kind=_lib.MASK_NOT, a=i)
        return i


class LazyMask(MaskBase):
    """masks.py:586-668: ``function`` evaluated on the data given at construction."""

    def __init__(self, function, cube=None, data=None, wcs=None):
        if cube is not None and (data is not None or wcs is not None):
            raise ValueError("Pass only cube or (data & wcs)")
        elif cube is not None:
            self._data, self._wcs = cube._data, cube._wcs
        elif data is not None:
            torch = _torch()
            self._data, self._wcs = _as_device(data, torch.float32), wcs
        else:
            raise ValueError("Either a cube or (data & wcs) is required.")
        self._function = function

    @property
    def shape(self):
        return tuple(self._data.shape)

    def _reference_tensor(self):
        return self._data

    def __getitem__(self, view):
        """masks.py:641-647: the same function on the sliced data (a view: no copy)."""
        return LazyMask(self._function, data=self._data[_slice3(view)], wcs=self._wcs)

    def _lower(self, low):
        if self._function is np.isfinite:
            return low.add(kind=_lib.MASK_FINITE, **low.data_fields(self._data))
        if self._function is np.isnan:
            # NaN test == not (x == x)
            i = low.add(kind=_lib.MASK_CMP_ARRAY, op=_lib.EQ, array_dtype=_lib.F32,
                        **low.data_fields(self._data), **low.array_fields(self._data))
            return low.add(kind=_lib.MASK_NOT, a=i)
        # arbitrary callable: it must accept a device tensor and return a boolean tensor
        torch = _torch()
        try:
            res = self._function(self._data)
        except Exception as exc:
            raise NotImplementedError(
                "LazyMask functions other than np.isfinite/np.isnan must accept a torch device "
                "tensor and return a boolean tensor (no CPU fallback): %r" % (exc,))
        if not isinstance(res, torch.Tensor) or tuple(res.shape) != tuple(self._data.shape):
            raise ValueError("Function did not return mask with correct shape")
        return low.add(kind=_lib.MASK_BOOL, array_dtype=_lib.U8,
                       **low.array_fields(res.to(device=low.cube.device, dtype=torch.uint8)))


class LazyComparisonMask(LazyMask):
    """masks.py:670-758"""

    def __init__(self, function, comparison_value, cube=None, data=None, wcs=None):
        if cube is not None and (data is not None or wcs is not None):
            raise ValueError("Pass only cube or (data & wcs)")
        elif cube is not None:
            self._data, self._wcs = cube._data, cube._wcs
        elif data is not None:
            torch = _torch()
            self._data, self._wcs = _as_device(data, torch.float32), wcs
        else:
            raise ValueError("Either a cube or (data & wcs) is required.")
        if function not in _OPS:
            raise NotImplementedError("comparison function must be one of operator.gt/ge/lt/le/eq/ne")
        if (hasattr(comparison_value, 'shape') and len(comparison_value.shape) > 0 and
                not is_broadcastable_and_smaller(tuple(self._data.shape), tuple(comparison_value.shape))):
            raise ValueError("The data and the comparison value cannot be broadcast to match shape")
        self._function = function
        self._comparison_value = comparison_value

    def __getitem__(self, view):
        """masks.py:735-745: sliced data; an array to compare with is sliced alongside (broadcast axes stay)."""
        cv = self._comparison_value
        if hasattr(cv, 'shape') and len(cv.shape) > 0:
            torch = _torch()
            t = cv if isinstance(cv, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cv, dtype=np.float64))
            cv = _slice_broadcast(t, view)
        return LazyComparisonMask(self._function, cv, data=self._data[_slice3(view)], wcs=self._wcs)

    def _lower(self, low):
        cv = self._comparison_value
        if hasattr(cv, 'shape') and len(cv.shape) > 0:
            torch = _torch()
            if isinstance(cv, torch.Tensor):
                t = cv.to(low.cube.device)
                if t.dtype not in (torch.float32, torch.float64):
                    t = t.to(torch.float64)
            else:
                t = torch.from_numpy(np.ascontiguousarray(cv, dtype=np.float64)).to(low.cube.device)
            dt = _lib.F32 if t.dtype == torch.float32 else _lib.F64
            return low.add(kind=_lib.MASK_CMP_ARRAY, op=_OPS[self._function], array_dtype=dt,
                           **low.data_fields(self._data), **low.array_fields(t))
        # scalar: reaches numpy as an np.float64 in the reference (spectral_cube.py:2248-2252)
        return low.add(kind=_lib.MASK_CMP_SCALAR, op=_OPS[self._function], value=float(cv),
                       **low.data_fields(self._data))


class FunctionMask(MaskBase):
    """masks.py:760-803: ``function(data, wcs, view)`` -> boolean array; evaluated on the cube
    the mask is applied to.  The function must work on a torch device tensor."""

    def __getitem__(self, view):
        return self                                              # evaluated on whatever data it is applied to

    def __init__(self, function):
        self._function = function

    def _lower(self, low):
        torch = _torch()
        res = self._function(low.cube, None, ())
        if tuple(res.shape) != tuple(low.cube.shape):
            raise ValueError("Function did not return mask with correct shape - expected "
                             "{0}, got {1}".format(tuple(low.cube.shape), tuple(res.shape)))
        if not isinstance(res, torch.Tensor):
            res = torch.from_numpy(np.ascontiguousarray(res))
        return low.add(kind=_lib.MASK_BOOL, array_dtype=_lib.U8,
                       **low.array_fields(res.to(device=low.cube.device, dtype=torch.uint8)))
