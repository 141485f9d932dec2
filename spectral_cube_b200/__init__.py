"""
spectral_cube_b200 -- B200-native implementation of spectral-cube's per-spaxel hot path
(moment maps under lazy masks, spectral/spatial smoothing, spectral interpolation,
reprojection) behind the reference's ``SpectralCube`` / ``DaskSpectralCube`` method surface.

Everything numerical runs in hand-written CUDA (sm_100a) from ``csrc/`` through the C ABI in
``include/sc_b200.h``; there is no CPU fallback.
"""
from .cube import (SpectralCube, DaskSpectralCube, BaseSpectralCube, VarianceWarning,
                   SmoothingWarning, BeamUnitsError, SpectralCubeWarning, SIGMA2FWHM,
                   VaryingResolutionSpectralCube, DaskVaryingResolutionSpectralCube, BeamWarning,
                   NonFiniteBeamsWarning)
from .beam import Beam, Beams, BeamError, NoBeamError, EllipticalGaussian2DKernel
from .masks import (MaskBase, InvertedMask, CompositeMask, BooleanArrayMask, LazyMask,
                    LazyComparisonMask, FunctionMask)
from .projection import Projection
from .wcs import CubeWCS
from .mosaic import mosaic_cubes, combine_headers, find_optimal_celestial_wcs
from .kernels import (Kernel1D, Kernel2D, Gaussian1DKernel, Gaussian2DKernel, Tophat2DKernel,
                      Box1DKernel, CustomKernel)

__all__ = ['SpectralCube', 'DaskSpectralCube', 'BaseSpectralCube', 'Projection', 'CubeWCS',
           'MaskBase', 'InvertedMask', 'CompositeMask', 'BooleanArrayMask', 'LazyMask',
           'LazyComparisonMask', 'FunctionMask', 'VarianceWarning', 'SmoothingWarning',
           'BeamUnitsError', 'SpectralCubeWarning', 'SIGMA2FWHM',
           'Kernel1D', 'Kernel2D', 'Gaussian1DKernel', 'Gaussian2DKernel', 'Tophat2DKernel',
           'Box1DKernel', 'CustomKernel', 'VaryingResolutionSpectralCube',
           'DaskVaryingResolutionSpectralCube', 'Beam', 'Beams', 'BeamError', 'NoBeamError', 'BeamWarning',
           'NonFiniteBeamsWarning', 'EllipticalGaussian2DKernel', 'mosaic_cubes', 'combine_headers',
           'find_optimal_celestial_wcs']
