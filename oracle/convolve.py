"""
oracle/convolve.py -- TEST INFRASTRUCTURE.  Restatement of ``astropy.convolution``.

The reference smooths by calling ``astropy.convolution.convolve(array, kernel,
normalize_kernel=True)`` (``spectral_cube/spectral_cube.py:2837`` per channel image,
``:3216-3222`` per spectrum; ``spectral_cube/dask_spectral_cube.py:912-914`` with a
``(n,1,1)`` kernel, ``:990-991`` per image).  astropy (>=6.1, ``pyproject.toml:26``) is
a third-party dependency that is absent here, so its *published* semantics are
restated (astropy docs, ``convolve``: defaults ``boundary='fill'``, ``fill_value=0``,
``nan_treatment='interpolate'``, ``preserve_nan=False``):

  * the kernel must have odd size along every axis and is applied flipped (true
    convolution);
  * the array is zero padded; padded zeros are *valid* samples;
  * ``out = sum_k K[k] v[.-k] [v not NaN] / sum_k K[k] [v not NaN]`` evaluated in
    float64; where the denominator is 0 the input value (NaN) is kept;
  * a float input dtype is restored on output (float32 in -> float32 out).

Kernel discretisation (``Gaussian1DKernel``, ``Gaussian2DKernel``, ``Tophat2DKernel``,
``Box1DKernel``): default size = ceil(8 sigma) (2 radius for the top-hat) rounded up to
odd, model sampled at integer offsets ('center' mode), Gaussians normalised to unit
sum (astropy PR 13299).  Pinned by the reference goldens in
``spectral_cube/tests/test_regrid.py:138-172`` and
``spectral_cube/tests/test_spectral_cube.py:2363-2421``.
"""
import math
import numpy as np


def _odd_ceil(value):
    i = int(math.ceil(value))
    return i + 1 if i % 2 == 0 else i


class Kernel(object):
    def __init__(self, array):
        self.array = np.asarray(array, dtype=np.float64)

    @property
    def shape(self):
        return self.array.shape


def Gaussian1DKernel(stddev, x_size=None):
    size = _odd_ceil(8 * stddev) if x_size is None else int(x_size)
    x = np.arange(-(size - 1) // 2, (size - 1) // 2 + 1, dtype=np.float64)
    a = np.exp(-0.5 * x ** 2 / stddev ** 2) / (np.sqrt(2 * np.pi) * stddev)
    return Kernel(a / a.sum())


def Gaussian2DKernel(x_stddev, y_stddev=None, theta=0.0, x_size=None, y_size=None):
    if y_stddev is None:
        y_stddev = x_stddev
    size = _odd_ceil(8 * max(x_stddev, y_stddev))
    x_size = size if x_size is None else int(x_size)
    y_size = x_size if y_size is None else int(y_size)
    x = np.arange(-(x_size - 1) // 2, (x_size - 1) // 2 + 1, dtype=np.float64)
    y = np.arange(-(y_size - 1) // 2, (y_size - 1) // 2 + 1, dtype=np.float64)
    yy, xx = np.meshgrid(y, x, indexing='ij')
    cost2, sint2, sin2t = np.cos(theta) ** 2, np.sin(theta) ** 2, np.sin(2.0 * theta)
    xs2, ys2 = x_stddev ** 2, y_stddev ** 2
    a = 0.5 * (cost2 / xs2 + sint2 / ys2)
    b = 0.5 * (sin2t / xs2 - sin2t / ys2)
    c = 0.5 * (sint2 / xs2 + cost2 / ys2)
    arr = np.exp(-(a * xx ** 2 + b * xx * yy + c * yy ** 2)) / (2 * np.pi * x_stddev * y_stddev)
    return Kernel(arr / arr.sum())


def Tophat2DKernel(radius):
    size = _odd_ceil(2 * radius)
    x = np.arange(-(size - 1) // 2, (size - 1) // 2 + 1, dtype=np.float64)
    yy, xx = np.meshgrid(x, x, indexing='ij')
    arr = np.where(xx ** 2 + yy ** 2 <= radius ** 2, 1.0 / (np.pi * radius ** 2), 0.0)
    return Kernel(arr / arr.sum())


def Box1DKernel(width):
    size = _odd_ceil(width)
    return Kernel(np.full(size, 1.0 / size))


def convolve(array, kernel, normalize_kernel=True):
    """N-d NaN-interpolating convolution with zero-fill boundary (see module docstring).

    Written as a sum over kernel taps of shifted, zero-padded copies, so it is exact
    for any dimensionality (1-d spectra, 2-d images, 3-d blocks with an (n,1,1)
    kernel).  float64 throughout; float input dtypes are restored at the end.
    """
    karr = kernel.array if hasattr(kernel, 'array') else np.asarray(kernel, dtype=np.float64)
    array = np.asarray(array)
    in_dtype = array.dtype
    if karr.ndim != array.ndim:
        raise Exception("array and kernel have differing number of dimensions.")
    if any(s % 2 == 0 for s in karr.shape):
        raise Exception("Kernel size must be odd in all axes.")
    ksum = karr.sum()
    if normalize_kernel and np.isclose(ksum, 0.0, atol=1e-8):
        raise ValueError("The kernel can't be normalized, because its sum is close to zero.")
    a = array.astype(np.float64)
    good = ~np.isnan(a)
    pad = [(s // 2, s // 2) for s in karr.shape]
    ap = np.pad(np.where(good, a, 0.0), pad)                 # NaN -> 0 contribution
    gp = np.pad(good.astype(np.float64), pad, constant_values=1.0)   # padding is valid
    top = np.zeros_like(a)
    bot = np.zeros_like(a)
    kflip = karr[tuple(slice(None, None, -1) for _ in karr.shape)]
    for idx in np.ndindex(*karr.shape):
        w = kflip[idx]
        sl = tuple(slice(i, i + n) for i, n in zip(idx, a.shape))
        top += w * ap[sl]
        bot += w * gp[sl]
    with np.errstate(invalid='ignore', divide='ignore'):
        out = np.where(bot == 0, a, top / bot)
    if not normalize_kernel:
        out = out * ksum
    if in_dtype.kind == 'f':
        out = out.astype(in_dtype)
    return out


def convolve_fft(array, kernel, normalize_kernel=True):
    """``astropy.convolution.convolve_fft`` with its defaults (``boundary='fill'``, ``fill_value=0``,
    ``nan_treatment='interpolate'``, ``psf_pad`` and ``fft_pad`` on, ``min_wt=0``, ``crop=True``): the numpy
    class's default in ``convolve_to`` (``spectral_cube.py:3336, 4128``).  Restated from astropy's published
    algorithm: NaN/inf -> 0, both arrays centred in a square power-of-two box at least array + kernel wide,
    product of the transforms; the same convolution of the validity weights (1 in the padding) divides the
    result, results whose weight is below 10 eps become 0.0; complex arithmetic, real float64 out.
    Mathematically this is ``convolve`` above except where a kernel window holds no valid sample:
    0.0 here, the NaN kept there."""
    karr = kernel.array if hasattr(kernel, 'array') else np.asarray(kernel, dtype=np.float64)
    array = np.array(array, dtype=complex)
    karr = np.array(karr, dtype=complex)
    if karr.ndim != array.ndim:
        raise Exception("array and kernel have differing number of dimensions.")
    nanmask = np.isnan(array) | np.isinf(array)
    array[nanmask] = 0
    ksum = karr.sum()                       # the kernel is always normalised inside; the scale returns at the end
    if abs(ksum) < 1e-8:
        raise ValueError("The kernel can't be normalized, because its sum is close to zero.")
    karr = karr / ksum
    kernel_scale = 1.0 if normalize_kernel else ksum
    ashape, kshape = np.array(array.shape), np.array(karr.shape)
    fsize = int(2 ** np.ceil(np.log2(np.max(ashape + kshape))))
    newshape = (fsize,) * array.ndim

    def centred(n, big):
        centre = big - (big + 1) // 2
        return slice(centre - n // 2, centre + (n + 1) // 2)

    asl = tuple(centred(n, fsize) for n in array.shape)
    ksl = tuple(centred(n, fsize) for n in karr.shape)
    bigarray = np.zeros(newshape, dtype=complex)
    bigarray[asl] = array
    bigkernel = np.zeros(newshape, dtype=complex)
    bigkernel[ksl] = karr
    kernfft = np.fft.fftn(np.fft.ifftshift(bigkernel))
    fftmult = np.fft.fftn(bigarray) * kernfft * kernel_scale
    bigimwt = np.ones(newshape, dtype=complex)
    bigimwt[asl] = 1.0 - nanmask
    bigimwt = np.real(np.fft.ifftn(np.fft.fftn(bigimwt) * kernfft))
    with np.errstate(divide='ignore', invalid='ignore'):
        rifft = np.fft.ifftn(fftmult) / bigimwt
    rifft[bigimwt < 10 * np.finfo(bigimwt.dtype).eps] = 0.0
    return rifft[asl].real
