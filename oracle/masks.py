"""
oracle/masks.py -- TEST INFRASTRUCTURE.  numpy restatement of the reference mask tree.

Follows ``spectral_cube/masks.py``:
  MaskBase.include/exclude/_filled/_flattened  :101-237   (operators :239-249)
  InvertedMask                                 :337-347
  CompositeMask._include                       :425-435
  BooleanArrayMask (broadcast via as_strided)  :457-584
  LazyMask._include                            :649-651
  LazyComparisonMask._include                  :721-733
  FunctionMask._include                        :786-790
WCS validation (``_validate_wcs``) is metadata only and is not restated.
"""
import numpy as np


def _broadcast_view(arr, shape, view):
    """Restates ``BooleanArrayMask._include`` (:555-557) / ``view_of_subset``: the stored
    array may be smaller than the cube but broadcastable to it."""
    return np.broadcast_to(arr, shape)[view]


class MaskBase(object):
    shape = None

    def include(self, data=None, view=()):
        return self._include(data=data, view=view)

    def exclude(self, data=None, view=()):
        return np.logical_not(self._include(data=data, view=view))       # :160-161

    def _filled(self, data, fill=np.nan, view=()):
        # :197-237 -- float dtype is preserved (float32 stays float32)
        dt = np.result_type(data.dtype, 0.0)
        sliced = data[view].astype(dt)
        ex = self.exclude(data=data, view=view)
        return np.ma.masked_array(sliced, mask=ex).filled(fill)

    def _flattened(self, data, view=()):
        return data[view][self.include(data=data, view=view)]            # :165-195

    def __and__(self, other):
        return CompositeMask(self, other, 'and')

    def __or__(self, other):
        return CompositeMask(self, other, 'or')

    def __xor__(self, other):
        return CompositeMask(self, other, 'xor')

    def __invert__(self):
        return InvertedMask(self)


class InvertedMask(MaskBase):
    def __init__(self, mask):
        self._mask = mask

    def _include(self, data=None, view=()):
        return np.logical_not(self._mask.include(data=data, view=view))


class CompositeMask(MaskBase):
    def __init__(self, mask1, mask2, operation='and'):
        self._mask1, self._mask2, self._operation = mask1, mask2, operation

    def _include(self, data=None, view=()):
        a = self._mask1._include(data=data, view=view)
        b = self._mask2._include(data=data, view=view)
        if self._operation == 'and':
            return np.bitwise_and(a, b)
        elif self._operation == 'or':
            return np.bitwise_or(a, b)
        elif self._operation == 'xor':
            return np.bitwise_xor(a, b)
        raise ValueError("Operation '{0}' not supported".format(self._operation))


class BooleanArrayMask(MaskBase):
    def __init__(self, mask, shape=None, include=True):
        mask = np.asarray(mask)
        self._mask_type = 'include' if include else 'exclude'
        self._mask = mask
        self.shape = tuple(shape) if shape is not None else mask.shape

    def _include(self, data=None, view=()):
        result = _broadcast_view(self._mask, self.shape, view)
        return result if self._mask_type == 'include' else np.logical_not(result)


class LazyMask(MaskBase):
    """``function`` evaluated on the data given at construction (NOT on the data passed
    to ``include``) -- masks.py:649-651; this is why a smoothed cube keeps a mask that
    still looks at the un-smoothed array (``spectral_cube.py:3043-3045``)."""

    def __init__(self, function, data):
        self._function = function
        self._data = data
        self.shape = data.shape

    def _include(self, data=None, view=()):
        return self._function(self._data[view])


class LazyComparisonMask(LazyMask):
    def __init__(self, function, comparison_value, data):
        self._function = function
        self._data = data
        self._comparison_value = comparison_value
        self.shape = data.shape

    def _include(self, data=None, view=()):
        cv = self._comparison_value
        if hasattr(cv, 'shape') and cv.shape:
            return self._function(self._data[view], _broadcast_view(cv, self._data.shape, view))
        return self._function(self._data[view], cv)


class FunctionMask(MaskBase):
    def __init__(self, function):
        self._function = function

    def _include(self, data=None, view=()):
        result = self._function(data, None, view)
        if result.shape != data[view].shape:
            raise ValueError("Function did not return mask with correct shape - expected "
                             "{0}, got {1}".format(data[view].shape, result.shape))
        return result
