"""
Convolution kernels with astropy's names and discretisation rules.

The reference's smoothing methods take ``astropy.convolution`` kernel objects
(``spectral_cube.py:2808-2842, 3186-3222``) and only ever read ``kernel.array`` (and check
that the kernel has no unit, :3212-3214).  astropy is not a dependency here, so these
classes reproduce the published defaults: size = ceil(8 sigma) (2 radius for the top-hat,
the width for a box) rounded up to odd, the model sampled at integer pixel offsets
(``mode='center'``), Gaussians and boxes normalised to unit sum.  Any object with an
``.array`` attribute (a real astropy kernel included) is accepted by the cube methods.
"""
import math

import numpy as np


def _odd_ceil(value):
    n = int(math.ceil(value))
    return n if n % 2 == 1 else n + 1


def _centred_axis(size):
    half = (size - 1) // 2
    return np.arange(-half, half + 1, dtype=np.float64)


class _Kernel(object):
    def __init__(self, array):
        self._array = np.asarray(array, dtype=np.float64)
        if any(n % 2 == 0 for n in self._array.shape):
            raise ValueError("Kernel size must be odd in all axes.")

    @property
    def array(self):
        return self._array

    @property
    def shape(self):
        return self._array.shape

    @property
    def dimension(self):
        return self._array.ndim


class Kernel1D(_Kernel):
    pass


class Kernel2D(_Kernel):
    pass


class CustomKernel(_Kernel):
    pass


class Gaussian1DKernel(Kernel1D):
    def __init__(self, stddev, x_size=None):
        size = _odd_ceil(8.0 * stddev) if x_size is None else int(x_size)
        x = _centred_axis(size)
        g = np.exp(-0.5 * (x / stddev) ** 2) / (math.sqrt(2.0 * math.pi) * stddev)
        super(Gaussian1DKernel, self).__init__(g / g.sum())
        self.stddev = stddev


class Box1DKernel(Kernel1D):
    def __init__(self, width):
        size = _odd_ceil(width)
        super(Box1DKernel, self).__init__(np.full(size, 1.0 / size))


class Gaussian2DKernel(Kernel2D):
    def __init__(self, x_stddev, y_stddev=None, theta=0.0, x_size=None, y_size=None):
        if y_stddev is None:
            y_stddev = x_stddev
        default = _odd_ceil(8.0 * max(x_stddev, y_stddev))
        x_size = default if x_size is None else int(x_size)
        y_size = x_size if y_size is None else int(y_size)
        yy, xx = np.meshgrid(_centred_axis(y_size), _centred_axis(x_size), indexing='ij')
        ct, st = math.cos(theta), math.sin(theta)
        a = 0.5 * (ct * ct / x_stddev ** 2 + st * st / y_stddev ** 2)
        b = 0.5 * math.sin(2.0 * theta) * (1.0 / x_stddev ** 2 - 1.0 / y_stddev ** 2)
        c = 0.5 * (st * st / x_stddev ** 2 + ct * ct / y_stddev ** 2)
        g = np.exp(-(a * xx * xx + b * xx * yy + c * yy * yy)) / (2.0 * math.pi * x_stddev * y_stddev)
        super(Gaussian2DKernel, self).__init__(g / g.sum())
        self.x_stddev, self.y_stddev, self.theta = x_stddev, y_stddev, theta


class Tophat2DKernel(Kernel2D):
    def __init__(self, radius):
        size = _odd_ceil(2.0 * radius)
        yy, xx = np.meshgrid(_centred_axis(size), _centred_axis(size), indexing='ij')
        disk = (xx * xx + yy * yy <= radius * radius).astype(np.float64)
        super(Tophat2DKernel, self).__init__(disk / disk.sum())
        self.radius = radius
