"""
oracle/ -- CPU restatement of the spectral-cube hot path.  TEST INFRASTRUCTURE ONLY.

This package re-states, in plain numpy/scipy, the reference algorithms that the
CUDA library in ``spectral_cube_b200/csrc`` replaces.  It is *not* a product code
path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``spectral_cube_b200``) never imports ``oracle`` and has no CPU fallback.

Why a restatement and not the reference itself: the reference
(radio-astro-tools/spectral-cube @ cb6969e) is pure Python on top of astropy,
dask, radio_beam and reproject, none of which is installable in the build
container or on the GPU box (no network, not in /opt/wheelhouse).  ``import
spectral_cube`` fails at ``spectral_cube/spectral_cube.py:14`` (dask) / ``:16``
(astropy).  The arithmetic on the path lives in

  * ``spectral_cube/_moments.py:30-202``            -> ``oracle/moments.py``
  * ``spectral_cube/masks.py:101-803``              -> ``oracle/masks.py``
  * ``astropy.convolution.convolve`` (astropy>=6.1, ``pyproject.toml:26``; call
    sites ``spectral_cube.py:2837, 3216-3222``; ``dask_spectral_cube.py:914, 991``)
                                                    -> ``oracle/convolve.py``
  * ``astropy.convolution.Gaussian1DKernel`` etc.   -> ``oracle/convolve.py``
  * ``numpy.interp`` / ``scipy.interpolate.interp1d`` (both importable here; call
    sites ``spectral_cube.py:3302-3310``, ``dask_spectral_cube.py:1346-1349``)
                                                    -> ``oracle/interp.py`` (calls the real functions)
  * ``astropy.convolution.convolve_fft`` (the numpy class's default in ``convolve_to``,
    ``spectral_cube.py:3336, 4128``)                 -> ``oracle/convolve.py:convolve_fft``
  * ``radio_beam.Beam.deconvolve / as_kernel / sr`` (radio-beam>=0.3.5, ``pyproject.toml:32``,
    not installed; call sites ``spectral_cube.py:3361-3378, 4186-4209``)
                                                    -> ``oracle/beam.py``
  * ``reproject.reproject_interp`` (reproject>=0.9.1, not installed)
                                                    -> ``oracle/reproject.py``
  * ``astropy.wcs`` (wcslib; linear spectral + TAN/SIN/CAR celestial)
                                                    -> ``oracle/wcs.py``

Pinning: ``tests/test_oracle_goldens.py`` checks this oracle
against every value-carrying golden the reference's own tests hold for the path:
``tests/test_moments.py:19-48`` (all 9 order x axis tables, three strategies, with
the ``> 4 K`` mask consistency test), ``tests/test_spectral_cube.py:2372-2383,
2410-2421`` (Gaussian2DKernel / Tophat2DKernel spatial smoothing, 7 decimals),
``tests/test_regrid.py:138-172`` (spectral smoothing of a delta), ``:234-248,
292-303, 318-345, 350-361`` (spectral interpolation); ``tests/test_convolve_to_host.py`` does the same for
``convolve_to``: ``tests/test_regrid.py:33-101``, ``tests/test_spectral_cube.py:2150-2225`` with the fixtures of
``conftest.py:590-660``.  ``reproject`` has no value
golden in the reference (``tests/test_regrid.py:99-135`` checks shape/WCS only) and
the package is absent: **bilinear reproject parity is unpinned** (restated from
the published algorithm; see ``oracle/reproject.py``).
"""
