"""Module path of the reference's ``spectral_cube.cube_utils`` for the functions on the hot path."""
from .mosaic import mosaic_cubes, combine_headers     # noqa: F401  (cube_utils.py:744-856)
from .cube import MEMORY_THRESHOLD                     # noqa: F401  (cube_utils.py:266-268)
